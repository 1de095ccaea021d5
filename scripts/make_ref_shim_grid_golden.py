"""Predictive grids from the reference's own predict SOURCE (bpl/base.py:74-148 and the per-class variants), run under
the jax / numpyro stand-ins of oracle/ref_shim.py on the inputs of tests/golden/grid_<model>.npz (posterior samples and
fixtures written by scripts/make_golden.py).  Writes tests/golden/refgrid_<model>.npz: same inputs, the reference's
grid [F, g, g] and outcome [F, 3]."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim
ref_shim.install()
import bpl

GOLD = os.path.join(ROOT, "tests", "golden")
CLS = {"dixon_coles": bpl.DixonColesMatchPredictor, "extended": bpl.ExtendedDixonColesMatchPredictor,
       "neutral": bpl.NeutralDixonColesMatchPredictor, "neutral_wc": bpl.NeutralDixonColesMatchPredictorWC}
for model, cls in CLS.items():
    z = np.load(os.path.join(GOLD, f"grid_{model}.npz"))
    s = {k[2:]: z[k] for k in z.files if k.startswith("s_")}
    fx = {k[2:]: z[k] for k in z.files if k.startswith("f_")}
    mg = int(z["max_goals"])
    m = cls()
    T = s["attack"].shape[1]
    m.teams = np.array([str(i) for i in range(T)])
    m._teams_dict = {str(i): i for i in range(T)}
    for k, v in s.items():
        setattr(m, k, torch.from_numpy(v.astype(np.float64)))
    h, a = [str(i) for i in fx["home_team"]], [str(i) for i in fx["away_team"]]
    if model == "neutral_wc":
        Cf = s["confederation_strength"].shape[1]
        m.conferences = np.array([str(i) for i in range(Cf)])
        m._conferences_dict = {str(i): i for i in range(Cf)}
        args = (h, a, [str(i) for i in fx["home_conf"]], [str(i) for i in fx["away_conf"]], fx["neutral_venue"].astype(np.int64))
    elif model == "neutral":
        args = (h, a, fx["neutral_venue"].astype(np.int64))
    else:
        args = (h, a)
    grid, _, _ = m.predict_score_grid_proba(*args, max_goals=mg)
    out = m.predict_outcome_proba(*args, max_goals=mg)
    # the API methods built on the grid, for the first fixture (bpl/base.py:248-348 and the neutral variants)
    t0, o0 = h[0], a[0]
    n = list(range(mg + 1))
    if model == "neutral_wc":
        tc, oc, nv0 = str(fx["home_conf"][0]), str(fx["away_conf"][0]), int(fx["neutral_venue"][0])
        api = dict(score_home=m.predict_score_n_proba(n, t0, o0, tc, oc, home=True, neutral_venue=nv0, max_goals=mg),
                   score_away=m.predict_score_n_proba(n, o0, t0, oc, tc, home=False, neutral_venue=nv0, max_goals=mg),
                   concede_home=m.predict_concede_n_proba(n, t0, o0, tc, oc, home=True, neutral_venue=nv0, max_goals=mg),
                   concede_away=m.predict_concede_n_proba(n, o0, t0, oc, tc, home=False, neutral_venue=nv0, max_goals=mg))
        ko = m.predict_outcome_proba(*args, knockout=True, max_goals=mg)
        pr = m.predict_score_proba(h, a, args[2], args[3], np.ones(len(h), dtype=int), np.zeros(len(h), dtype=int), args[4])
    elif model == "neutral":
        nv0 = int(fx["neutral_venue"][0])
        api = dict(score_home=m.predict_score_n_proba(n, t0, o0, home=True, neutral_venue=nv0, max_goals=mg),
                   score_away=m.predict_score_n_proba(n, o0, t0, home=False, neutral_venue=nv0, max_goals=mg),
                   concede_home=m.predict_concede_n_proba(n, t0, o0, home=True, neutral_venue=nv0, max_goals=mg),
                   concede_away=m.predict_concede_n_proba(n, o0, t0, home=False, neutral_venue=nv0, max_goals=mg))
        ko = m.predict_outcome_proba(*args, knockout=True, max_goals=mg)
        pr = m.predict_score_proba(h, a, np.ones(len(h), dtype=int), np.zeros(len(h), dtype=int), args[2])
    else:
        api = dict(score_home=m.predict_score_n_proba(n, t0, o0, home=True, max_goals=mg),
                   score_away=m.predict_score_n_proba(n, o0, t0, home=False, max_goals=mg),
                   concede_home=m.predict_concede_n_proba(n, t0, o0, home=True, max_goals=mg),
                   concede_away=m.predict_concede_n_proba(n, o0, t0, home=False, max_goals=mg))
        ko = None
        pr = m.predict_score_proba(h, a, np.ones(len(h), dtype=int), np.zeros(len(h), dtype=int))
    api = {"api_" + k: np.asarray(v) for k, v in api.items()}
    api["api_score_1_0"] = np.asarray(pr)
    if ko is not None:
        api["api_knockout"] = np.stack([np.asarray(ko["home_win"]), np.asarray(ko["away_win"])], 1)
    grid = np.asarray(grid)
    outcome = np.stack([np.asarray(out["home_win"]), np.asarray(out["draw"]), np.asarray(out["away_win"])], 1)
    np.savez_compressed(os.path.join(GOLD, f"refgrid_{model}.npz"), grid=grid, outcome=outcome, **api,
                        **{k: z[k] for k in z.files if k not in ("grid", "outcome")})
    print(f"{model:12s} max |reference - oracle| grid {np.abs(grid - z['grid']).max():.2e} outcome {np.abs(outcome - z['outcome']).max():.2e}")
