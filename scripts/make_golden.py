#!/usr/bin/env python
"""Writes tests/golden/*.npz: inputs + float64 oracle outputs for small seeded problems.

The reference itself cannot be run in this image (no jax / numpyro), so these vectors come from the
oracle restatement (oracle/models.py, oracle/predict.py) -- they pin the oracle against silent change and
give the GPU tests a committed, oracle-independent-at-runtime target.  Re-run: python scripts/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import datasets, models as om, predict as op  # noqa: E402
from tests import helpers as H  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

ARR_KEYS = ("home_team", "away_team", "home_goals", "away_goals", "weights", "neutral_venue", "home_conf", "away_conf",
            "covariates", "gameweek")


def save_density(name, arr, C=8, radius=1.0, seed=0):
    d = H.to_oracle(arr)
    D = om.num_params(arr.model, arr.num_teams, arr.num_covariates, arr.num_conferences, arr.num_gameweeks)
    theta = H.random_theta(D, C, seed=seed, radius=radius).astype(np.float32).astype(np.float64)
    lp, g, cc = om.log_density_and_grad(d, theta)
    rec = {k: getattr(arr, k) for k in ARR_KEYS if getattr(arr, k, None) is not None}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), model=arr.model, num_teams=arr.num_teams,
                        num_conferences=arr.num_conferences, num_gameweeks=arr.num_gameweeks,
                        theta=theta.astype(np.float32), lp=lp, grad=g, corr_coef=cc, **rec)
    print(name, "D", D, "lp[0]", lp[0])


save_density("dc_dummy", H.from_training_data("dixon_coles", datasets.dummy_data()))
save_density("ext_timed", H.from_training_data("extended", datasets.timed_dummy_data(), epsilon=1.0))
save_density("ext_small_cov", H.small_problem("extended", seed=7, T=9, M=120, K=3, weighted=True), radius=1.5)
save_density("neutral_dummy", H.from_training_data("neutral", datasets.neutral_dummy_data(), epsilon=0.5))
save_density("wc_dummy", H.from_training_data("neutral_wc", datasets.neutral_dummy_data(), epsilon=0.2))
save_density("wc_small_multiconf", H.small_problem("neutral_wc", seed=9, T=11, M=300, K=2, multi_conf=True), radius=2.0)

# predictive grids
rng = np.random.default_rng(11)
for model in ("dixon_coles", "extended", "neutral", "neutral_wc"):
    S, T, F, Cf, mg = 96, 8, 12, 3, 10
    s = {"attack": rng.normal(0, 0.3, (S, T)), "defence": rng.normal(0, 0.3, (S, T)),
         "corr_coef": rng.uniform(-0.15, 0.15, S)}
    if model == "dixon_coles":
        s["home_advantage"] = rng.normal(0.25, 0.1, S)
    elif model == "extended":
        s["home_advantage"] = rng.normal(0.25, 0.1, (S, T))
    else:
        for k, m in (("home_attack", 0.1), ("away_attack", -0.1), ("home_defence", 0.1), ("away_defence", -0.1)):
            s[k] = rng.normal(m, 0.1, (S, T))
        if model == "neutral_wc":
            s["confederation_strength"] = rng.normal(0, 0.2, (S, Cf))
    s = {k: v.astype(np.float32) for k, v in s.items()}
    h = rng.integers(0, T, F)
    a = (h + rng.integers(1, T, F)) % T
    fx = {"home_team": h.astype(np.uint16), "away_team": a.astype(np.uint16)}
    kw = {}
    if model in ("neutral", "neutral_wc"):
        fx["neutral_venue"] = (rng.random(F) < 0.4).astype(np.uint8)
        kw["neutral_venue"] = fx["neutral_venue"]
    if model == "neutral_wc":
        fx["home_conf"] = rng.integers(0, Cf, F).astype(np.uint8)
        fx["away_conf"] = rng.integers(0, Cf, F).astype(np.uint8)
        kw["home_conf"], kw["away_conf"] = fx["home_conf"], fx["away_conf"]
    grid, _, _ = op.predict_score_grid_proba(model, s, h, a, mg, **kw)
    out = op.predict_outcome_proba(model, s, h, a, mg, **kw)
    np.savez_compressed(os.path.join(OUT, f"grid_{model}.npz"), model=model, max_goals=mg, grid=grid,
                        outcome=np.stack([out["home_win"], out["draw"], out["away_win"]], 1),
                        **{"s_" + k: v for k, v in s.items()}, **{"f_" + k: v for k, v in fx.items()})
    print("grid", model, grid.sum(axis=(1, 2))[:3])
