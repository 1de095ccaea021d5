"""Closed-form value + gradient of the static Dixon-Coles-family log-densities.

TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by ``bpl_next_b200``.

``oracle/models.py`` restates the reference line by line and lets autograd differentiate it; that is the parity
oracle.  This file is the *fast* CPU restatement that ``bench.py`` times as ``cpu_baseline`` and in the
``--impl reference`` arm: the same forward arithmetic (``bpl/dixon_coles.py:39-84``, ``bpl/extended_dixon_coles.py:78-248``,
``bpl/neutral_dixon_coles.py:102-283``, ``bpl/neutral_dixon_coles_WC.py:83-232``, ``bpl/_util.py:17-93``) with the
hand-derived backward pass a fused XLA:CPU executable of ``jax.value_and_grad(potential_fn)`` amounts to -- one pass over
``[chains, matches]`` float32 arrays, scatter-adds per team -- vectorised over a block of chains, all host threads.
It is checked against the autograd oracle in ``tests/test_oracle.py`` (float64: 1e-9 relative).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import models as om

_EFFECTS = ("home_attack", "away_attack", "home_defence", "away_defence")


class ClosedForm:
    """Static per-problem tensors (indices, masks, weights) prepared once; ``__call__(theta [C, D])``."""

    def __init__(self, d: om.MatchData, dtype=torch.float32):
        if d.model not in ("dixon_coles", "extended", "neutral", "neutral_wc"):
            raise ValueError(f"closed form covers the four static models, not {d.model!r}")
        self.d, self.dt = d, dtype
        T, K = d.num_teams, d.num_covariates
        self.T, self.K = T, K
        self.Cf = d.num_conferences if d.model == "neutral_wc" else 0
        self.lay = om.layout_offsets(om.site_layout(d.model, T, K, self.Cf))
        self.D = self.lay["__D__"][0]
        self.h, self.a = om._idx(d.home_team), om._idx(d.away_team)
        hg, ag = np.asarray(d.home_goals), np.asarray(d.away_goals)
        self.yh = torch.as_tensor(hg, dtype=dtype)
        self.ya = torch.as_tensor(ag, dtype=dtype)
        M = len(hg)
        self.w = torch.ones(M, dtype=dtype) if d.weights is None else torch.as_tensor(np.asarray(d.weights), dtype=dtype)
        self.const = float(-(self.w.double() * (torch.lgamma(self.yh.double() + 1) + torch.lgamma(self.ya.double() + 1))).sum())
        self.i00 = om._idx(np.nonzero((hg == 0) & (ag == 0))[0])
        self.i10 = om._idx(np.nonzero((hg == 1) & (ag == 0))[0])
        self.i01 = om._idx(np.nonzero((hg == 0) & (ag == 1))[0])
        self.w11 = float(self.w[om._idx(np.nonzero((hg == 1) & (ag == 1))[0])].double().sum())
        self.neutral = d.model in ("neutral", "neutral_wc")
        if self.neutral:
            self.n = torch.as_tensor(1 - np.asarray(d.neutral_venue, dtype=np.int64), dtype=dtype)
        if self.Cf:
            self.hc, self.ac = om._idx(d.home_conf), om._idx(d.away_conf)
        self.Xs = None
        if d.covariates is not None:
            Xs = d.covariates if d.covariates_prestandardised else om.standardise_covariates(d.covariates)
            self.Xs = torch.as_tensor(np.asarray(Xs, dtype=np.float64), dtype=dtype)
        self.clip = d.model == "extended"

    # -- small helpers ----------------------------------------------------------------------------
    def _site(self, theta, name):
        o, shape, _ = self.lay[name]
        n = int(np.prod(shape)) if shape else 1
        x = theta[:, o:o + n]
        return x[:, 0] if shape == () else x

    def _put(self, grad, name, val):
        o, shape, _ = self.lay[name]
        n = int(np.prod(shape)) if shape else 1
        grad[:, o:o + n] = val.reshape(grad.shape[0], n)

    @staticmethod
    def _scatter(C, T, idx, val):
        out = torch.zeros((C, T), dtype=val.dtype)
        out.index_add_(1, idx, val)
        return out

    def __call__(self, theta):
        dt = self.dt
        th = torch.as_tensor(np.asarray(theta), dtype=dt)
        C = th.shape[0]
        T, K, model = self.T, self.K, self.d.model
        grad = torch.zeros((C, self.D), dtype=dt)
        lp = torch.full((C,), self.const, dtype=dt)
        LS2PI, LOG2 = om.LOG_SQRT_2PI, om.LOG2

        def normal(name, loc, scale):  # prior of a real site: value added to lp, returns d lp / d x
            nonlocal lp
            x = self._site(th, name)
            z = (x - loc) / scale
            t = -0.5 * z * z - math.log(scale) - LS2PI
            lp = lp + (t if t.dim() == 1 else t.sum(-1))
            return -z / scale

        def halfnormal(name, scale):  # HalfNormal(scale) on exp(x) + Jacobian x
            nonlocal lp
            x = self._site(th, name)
            s = torch.exp(x)
            z = s / scale
            lp = lp + (-0.5 * z * z - math.log(scale) - LS2PI + LOG2 + x)
            return s, 1.0 - z * z

        def beta(name, a, b):  # Beta(a, b) on clipped sigmoid(x) + Jacobian
            nonlocal lp
            x = self._site(th, name)
            fi = torch.finfo(dt)
            u = torch.clamp(torch.sigmoid(x), min=fi.tiny, max=1.0 - fi.eps)
            sp = torch.nn.functional.softplus
            lognorm = math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b)
            lp = lp + ((a - 1.0) * torch.log(u) + (b - 1.0) * torch.log1p(-u) - lognorm - sp(x) - sp(-x))
            return u, a * (1.0 - u) - b * u

        # ---- priors on the scalar sites; constrained values --------------------------------------
        eff_names = []
        if model == "dixon_coles":
            g_ha = normal("home_advantage", 0.1, 0.2)
            g_md = normal("mean_defence", 0.0, 1.0)
            sa, g_lsa = halfnormal("std_attack", 1.0)
            sd, g_lsd = halfnormal("std_defence", 1.0)
            za, zd = self._site(th, "attack_decentered"), self._site(th, "defence_decentered")
        else:
            pri = 0.5 if self.neutral else 1.0
            g_md = normal("mean_defence", 0.0, 1.0)
            sa, g_lsa = halfnormal("std_attack", pri)
            sd, g_lsd = halfnormal("std_defence", pri)
            za, zd = self._site(th, "standardised_attack"), self._site(th, "standardised_defence")
            if model == "extended":
                eff_names = [("home_advantage", "mean_home_advantage", "std_home_advantage", 0.1)]
            else:
                eff_names = [(nm, "mean_" + nm, "std_" + nm, 0.1 if nm in ("home_attack", "home_defence") else -0.1)
                             for nm in _EFFECTS]
        effs = {}
        for nm, mean_site, std_site, loc in eff_names:
            g_mu = normal(mean_site, loc, 0.2)
            s, g_ls = halfnormal(std_site, 1.0)
            dec = self._site(th, nm + "_decentered")
            lp = lp + (-0.5 * dec * dec - LS2PI).sum(-1)
            effs[nm] = dict(mu=self._site(th, mean_site), s=s, dec=dec, g_mu=g_mu, g_ls=g_ls,
                            val=self._site(th, mean_site)[:, None] + s[:, None] * dec, mean_site=mean_site, std_site=std_site)
        a_mean = torch.zeros((C, 1), dtype=dt)
        d_mean = self._site(th, "mean_defence")[:, None]
        if K:
            ba, bd = self._site(th, "attack_coefficients"), self._site(th, "defence_coefficients")
            lp = lp + (-0.5 * ba * ba - LS2PI).sum(-1) + (-0.5 * bd * bd - LS2PI).sum(-1)
            a_mean = ba @ self.Xs.T
            d_mean = d_mean + bd @ self.Xs.T
        if model == "dixon_coles":
            lp = lp + (-0.5 * za * za - LS2PI).sum(-1) + (-0.5 * zd * zd - LS2PI).sum(-1)
            p_za, p_zd = -za, -zd
        else:  # u ~ Beta(2,4), rho = 2u-1, za ~ N(0,1), zd ~ N(rho za, sqrt(1-rho^2))
            u, g_u = beta("u", 2.0, 4.0)
            rho = 2.0 * u - 1.0
            s2 = 1.0 - rho * rho
            e = zd - rho[:, None] * za
            lp = lp + (-0.5 * za * za - LS2PI).sum(-1) + (-0.5 * e * e / s2[:, None] - LS2PI).sum(-1) - 0.5 * T * torch.log(s2)
            p_zd = -e / s2[:, None]
            p_za = -za + rho[:, None] * e / s2[:, None]
            d_rho = (e * za / s2[:, None]).sum(-1) - rho * (e * e).sum(-1) / (s2 * s2) + T * rho / s2
            g_u = g_u + d_rho * 2.0 * u * (1.0 - u)
        attack = a_mean + za * sa[:, None]
        defence = d_mean + zd * sd[:, None]
        raw, g_raw = beta("corr_coef_raw", 2.0, 2.0)

        # ---- rates ----------------------------------------------------------------------------------
        h, a = self.h, self.a
        eta_h = attack[:, h] - defence[:, a]
        eta_a = attack[:, a] - defence[:, h]
        if model == "dixon_coles":
            eta_h = eta_h + self._site(th, "home_advantage")[:, None]
        elif model == "extended":
            eta_h = eta_h + effs["home_advantage"]["val"][:, h]
        else:
            n = self.n
            eta_h = eta_h + n * (effs["home_attack"]["val"][:, h] - effs["away_defence"]["val"][:, a])
            eta_a = eta_a + n * (effs["away_attack"]["val"][:, a] - effs["home_defence"]["val"][:, h])
        if self.Cf:
            conf = self._site(th, "confederation_strength_decentered")
            lp = lp + (-0.5 * conf * conf - LS2PI).sum(-1)
            dcf = conf[:, self.hc] - conf[:, self.ac]
            eta_h = eta_h + dcf
            eta_a = eta_a - dcf
        lam_h, lam_a = torch.exp(eta_h), torch.exp(eta_a)
        if self.clip:
            free_h, free_a = lam_h < 15.0, lam_a < 15.0  # d min(x, 15) / dx
            lam_h = torch.clamp(lam_h, max=15.0)
            lam_a = torch.clamp(lam_a, max=15.0)
        w, yh, ya = self.w, self.yh, self.ya
        lp = lp + (w * (torch.log(lam_h) * yh - lam_h)).sum(-1) + (w * (torch.log(lam_a) * ya - lam_a)).sum(-1)
        r_h = w * (yh - lam_h)  # d / d eta (log-rate) of the Poisson part
        r_a = w * (ya - lam_a)

        # ---- bounds, corr_coef (bpl/_util.py:17-31) -------------------------------------------------
        mh, ih = lam_h.max(-1)
        ma, ia = lam_a.max(-1)
        prod = lam_h * lam_a
        mp, ip = prod.max(-1)
        Lam = torch.maximum(mh, ma)
        LB = -1.0 / Lam
        UB = torch.clamp(1.0 / mp, max=1.0)
        cc = LB + raw * (UB - LB)
        c = cc[:, None]

        # ---- tau terms (bpl/_util.py:54-91) ----------------------------------------------------------
        Gc = torch.zeros(C, dtype=dt)
        for idx, kind in ((self.i00, 0), (self.i10, 1), (self.i01, 2)):
            if idx.numel() == 0:
                continue
            wi = w[idx]
            if kind == 0:
                t = lam_h[:, idx] * lam_a[:, idx]
                tau = torch.clamp(1.0 - c * t, min=0.0)
                q = wi * t / tau
                Gc = Gc - q.sum(-1)
                r_h[:, idx] -= c * q
                r_a[:, idx] -= c * q
            elif kind == 1:  # home 1, away 0: tau = 1 + c lam_a
                t = lam_a[:, idx]
                tau = torch.clamp(1.0 + c * t, min=0.0)
                q = wi * t / tau
                Gc = Gc + q.sum(-1)
                r_a[:, idx] += c * q
            else:  # home 0, away 1: tau = 1 + c lam_h
                t = lam_h[:, idx]
                tau = torch.clamp(1.0 + c * t, min=0.0)
                q = wi * t / tau
                Gc = Gc + q.sum(-1)
                r_h[:, idx] += c * q
            lp = lp + (wi * torch.log(tau)).sum(-1)
        t11 = torch.clamp(1.0 - cc, min=0.0)
        if self.w11 != 0.0:
            lp = lp + self.w11 * torch.log(t11)
            Gc = Gc - self.w11 / t11
        # arg-max fix-up (Appendix B.3)
        rows = torch.arange(C)
        use_h = mh >= ma
        fl = Gc * (1.0 - raw) / Lam
        r_h[rows, ih] += torch.where(use_h, fl, torch.zeros_like(fl))
        r_a[rows, ia] += torch.where(use_h, torch.zeros_like(fl), fl)
        fu = torch.where(mp > 1.0, -Gc * raw / mp, torch.zeros_like(mp))
        r_h[rows, ip] += fu
        r_a[rows, ip] += fu
        if self.clip:  # no gradient through a clipped rate
            r_h = r_h * free_h
            r_a = r_a * free_a
        g_raw = g_raw + Gc * raw * (1.0 - raw) * (UB - LB)

        # ---- scatter to teams, chain rule (Appendix B.4) -------------------------------------------
        sc = self._scatter
        g_att = sc(C, T, h, r_h) + sc(C, T, a, r_a)
        g_def = -(sc(C, T, a, r_h) + sc(C, T, h, r_a))
        self._put(grad, "corr_coef_raw", g_raw)
        if model == "dixon_coles":
            self._put(grad, "home_advantage", g_ha + r_h.sum(-1))
        elif model == "extended":
            effs["home_advantage"]["g"] = sc(C, T, h, r_h)
        else:
            nh, na = self.n * r_h, self.n * r_a
            effs["home_attack"]["g"] = sc(C, T, h, nh)
            effs["away_defence"]["g"] = -sc(C, T, a, nh)
            effs["away_attack"]["g"] = sc(C, T, a, na)
            effs["home_defence"]["g"] = -sc(C, T, h, na)
        for nm, e_ in effs.items():
            g = e_["g"]
            self._put(grad, nm + "_decentered", e_["s"][:, None] * g - e_["dec"])
            self._put(grad, e_["mean_site"], e_["g_mu"] + g.sum(-1))
            self._put(grad, e_["std_site"], e_["g_ls"] + e_["s"] * (e_["dec"] * g).sum(-1))
        if self.Cf:
            dr = r_h - r_a
            self._put(grad, "confederation_strength_decentered", sc(C, self.Cf, self.hc, dr) - sc(C, self.Cf, self.ac, dr) - conf)
        za_name = "attack_decentered" if model == "dixon_coles" else "standardised_attack"
        zd_name = "defence_decentered" if model == "dixon_coles" else "standardised_defence"
        self._put(grad, za_name, sa[:, None] * g_att + p_za)
        self._put(grad, zd_name, sd[:, None] * g_def + p_zd)
        self._put(grad, "std_attack", g_lsa + sa * (za * g_att).sum(-1))
        self._put(grad, "std_defence", g_lsd + sd * (zd * g_def).sum(-1))
        self._put(grad, "mean_defence", g_md + g_def.sum(-1))
        if model != "dixon_coles":
            self._put(grad, "u", g_u)
        if K:
            self._put(grad, "attack_coefficients", g_att @ self.Xs - ba)
            self._put(grad, "defence_coefficients", g_def @ self.Xs - bd)
        return lp.numpy(), grad.numpy(), cc.numpy()


def log_density_and_grad(d: om.MatchData, theta, dtype=torch.float32):
    return ClosedForm(d, dtype)(theta)
