"""Restatement of the posterior-predictive path: expected goals, score proba, grid, outcome.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``: pinned to the reference's predict source run under stand-ins for jax / numpyro).

Literal numpy restatement (float64 by default; ``dtype=np.float32`` reproduces the reference's
working precision) of
  * ``_calculate_expected_goals``: ``dixon_coles.py:126-137``, ``extended_dixon_coles.py:335-359``,
    ``neutral_dixon_coles.py:399-444``, ``neutral_dixon_coles_WC.py:363-422``
  * ``predict_score_proba``: ``dixon_coles.py:139-163``, ``extended...:361-399``,
    ``neutral...:446-488``, ``..._WC.py:424-474``
  * ``predict_score_grid_proba`` / ``predict_outcome_proba``: ``base.py:74-148``,
    ``neutral...:562-659``, ``..._WC.py:548-670``.
``samples`` is a dict of posterior arrays with the attribute names the reference's ``fit`` stores.
"""
from __future__ import annotations

import numpy as np
from scipy.special import gammaln


def expected_goals(model, samples, home, away, home_conf=None, away_conf=None, neutral_venue=None, dtype=np.float64):
    s = {k: np.asarray(v, dtype=dtype) for k, v in samples.items() if v is not None}
    h, a = np.asarray(home, dtype=np.int64), np.asarray(away, dtype=np.int64)
    att, dfn = s["attack"], s["defence"]
    if model == "dixon_coles":
        lh = np.exp(att[:, h] - dfn[:, a] + s["home_advantage"][:, None])
        la = np.exp(att[:, a] - dfn[:, h])
    elif model == "extended":
        lh = np.exp(att[:, h] - dfn[:, a] + s["home_advantage"][:, h])
        la = np.exp(att[:, a] - dfn[:, h])
    elif model in ("neutral", "neutral_wc"):
        n = (1 - np.asarray(neutral_venue, dtype=np.int64)).astype(dtype)
        eh = att[:, h] - dfn[:, a]
        ea = att[:, a] - dfn[:, h]
        if model == "neutral_wc":
            hc, ac = np.asarray(home_conf, dtype=np.int64), np.asarray(away_conf, dtype=np.int64)
            cs = s["confederation_strength"]
            eh = eh + cs[:, hc] - cs[:, ac]
            ea = ea + cs[:, ac] - cs[:, hc]
        eh = eh + n * s["home_attack"][:, h] - n * s["away_defence"][:, a]
        ea = ea + n * s["away_attack"][:, a] - n * s["home_defence"][:, h]
        lh, la = np.exp(eh), np.exp(ea)
    else:
        raise ValueError(model)
    return lh, la


def correlation_term(hg, ag, lh, la, corr_coef, tol=0.0):
    """``_util.py:35-93`` with ``weights=None`` -> ones."""
    hg, ag = np.asarray(hg), np.asarray(ag)
    corr = np.zeros_like(lh)
    c = corr_coef[..., None]
    with np.errstate(divide="ignore"):
        m = (hg == 0) & (ag == 0)
        corr[..., m] = np.log(np.clip(1.0 - c * lh[..., m] * la[..., m], tol, None))
        m = (hg == 1) & (ag == 0)
        corr[..., m] = np.log(np.clip(1.0 + c * la[..., m], tol, None))
        m = (hg == 0) & (ag == 1)
        corr[..., m] = np.log(np.clip(1.0 + c * lh[..., m], tol, None))
        m = (hg == 1) & (ag == 1)
        corr[..., m] = np.log(np.clip(np.broadcast_to(1.0 - c, lh[..., m].shape), tol, None))
    return corr


def _pois_lp(k, rate):
    return np.log(rate) * k - gammaln(k + 1.0) - rate


def predict_score_proba(model, samples, home, away, hg, ag, dtype=np.float64, **fx):
    lh, la = expected_goals(model, samples, home, away, dtype=dtype, **fx)
    cc = np.asarray(samples["corr_coef"], dtype=dtype)
    hg = np.asarray(hg).astype(dtype); ag = np.asarray(ag).astype(dtype)
    corr = correlation_term(hg, ag, lh, la, cc)
    if model in ("dixon_coles", "extended"):
        p = np.exp(corr) * np.exp(_pois_lp(hg, lh)) * np.exp(_pois_lp(ag, la))
    else:
        p = np.exp(corr + _pois_lp(hg, lh) + _pois_lp(ag, la))
    return p.mean(axis=0)


def predict_score_grid_proba(model, samples, home, away, max_goals=15, dtype=np.float64, chunk=64, **fx):
    """``base.py:74-111`` -- evaluated per chunk of fixtures so the [S, F*g^2] temporaries fit."""
    home, away = np.asarray(home), np.asarray(away)
    g = max_goals + 1
    n_goals = np.arange(0, g)
    HG, AG = np.meshgrid(n_goals, n_goals, indexing="ij")
    out = np.empty((len(home), g, g), dtype=dtype)
    for lo in range(0, len(home), chunk):
        hi = min(lo + chunk, len(home))
        nf = hi - lo
        hgf = np.tile(HG.reshape(g * g), nf)
        agf = np.tile(AG.reshape(g * g), nf)
        rep = {k: np.repeat(np.asarray(v)[lo:hi], g * g) for k, v in fx.items() if v is not None}
        p = predict_score_proba(model, samples, np.repeat(home[lo:hi], g * g), np.repeat(away[lo:hi], g * g),
                                hgf, agf, dtype=dtype, **rep)
        out[lo:hi] = p.reshape(nf, g, g)
    return out, HG, AG


def predict_outcome_proba(model, samples, home, away, max_goals=15, knockout=False, dtype=np.float64, **fx):
    probs, HG, AG = predict_score_grid_proba(model, samples, home, away, max_goals, dtype=dtype, **fx)
    hw = probs[:, HG > AG].sum(axis=-1)
    dr = probs[:, HG == AG].sum(axis=-1)
    aw = probs[:, HG < AG].sum(axis=-1)
    if knockout:  # neutral_dixon_coles.py:650-653
        norm = hw + aw
        return {"home_win": hw / norm, "away_win": aw / norm}
    return {"home_win": hw, "draw": dr, "away_win": aw}
