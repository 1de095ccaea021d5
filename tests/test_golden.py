"""Committed golden vectors (tests/golden/*.npz, written by scripts/make_golden.py from the float64 oracle).

CPU: the oracle and the double-precision plan walker still reproduce them.  GPU: K1 / K3 through the C ABI
match them within BASELINE.json's tolerances (lp 1e-5 rel, gradient 1e-4 rel, grids 1e-6 abs)."""
import glob
import os

import numpy as np
import pytest

from bpl_next_b200 import data as bdata
from oracle import models as om, predict as op
from tests import helpers as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DENSITY = sorted(f for f in glob.glob(os.path.join(GOLD, "*.npz")) if not os.path.basename(f).startswith("grid_"))
GRIDS = sorted(glob.glob(os.path.join(GOLD, "grid_*.npz")))


def _arrays(z):
    kw = {k: z[k] for k in ("weights", "neutral_venue", "home_conf", "away_conf", "covariates", "gameweek") if k in z.files}
    return bdata.MatchArrays(model=str(z["model"]), num_teams=int(z["num_teams"]), home_team=z["home_team"],
                             away_team=z["away_team"], home_goals=z["home_goals"], away_goals=z["away_goals"],
                             num_conferences=int(z["num_conferences"]), num_gameweeks=int(z["num_gameweeks"]), **kw)


def test_golden_files_present():
    assert len(DENSITY) >= 6 and len(GRIDS) == 4


@pytest.mark.parametrize("path", DENSITY, ids=os.path.basename)
def test_oracle_and_plan_reproduce_golden(path):
    z = np.load(path)
    arr = _arrays(z)
    theta = z["theta"].astype(np.float64)
    lp, g, cc = om.log_density_and_grad(H.to_oracle(arr), theta)
    np.testing.assert_allclose(lp, z["lp"], rtol=1e-12)
    np.testing.assert_allclose(g, z["grad"], rtol=1e-9, atol=1e-9 * np.abs(z["grad"]).max())
    lp_p, g_p, cc_p = H.plancheck_eval(arr, theta)
    np.testing.assert_allclose(lp_p, z["lp"], rtol=1e-7)
    scale = np.abs(z["grad"]).max(axis=1, keepdims=True)
    np.testing.assert_allclose(g_p / scale, z["grad"] / scale, rtol=0, atol=2e-7)
    np.testing.assert_allclose(cc_p, z["corr_coef"], rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("path", DENSITY, ids=os.path.basename)
def test_k1_matches_golden(path):
    from bpl_next_b200 import Problem

    z = np.load(path)
    p = Problem(_arrays(z))
    lp, grad, cc = p.logdensity_host(z["theta"])
    np.testing.assert_allclose(lp, z["lp"], rtol=1e-5)
    scale = np.abs(z["grad"]).max(axis=1, keepdims=True)
    assert (np.abs(grad - z["grad"]) / scale).max() < 1e-4
    np.testing.assert_allclose(cc, z["corr_coef"], rtol=1e-4, atol=1e-6)


def _grid_inputs(z):
    s = {k[2:]: z[k] for k in z.files if k.startswith("s_")}
    fx = {k[2:]: z[k] for k in z.files if k.startswith("f_")}
    return str(z["model"]), s, fx, int(z["max_goals"])


@pytest.mark.parametrize("path", GRIDS, ids=os.path.basename)
def test_predict_oracle_reproduces_golden(path):
    z = np.load(path)
    model, s, fx, mg = _grid_inputs(z)
    kw = {k: fx[k] for k in ("home_conf", "away_conf", "neutral_venue") if k in fx}
    grid, _, _ = op.predict_score_grid_proba(model, s, fx["home_team"], fx["away_team"], mg, **kw)
    np.testing.assert_allclose(grid, z["grid"], rtol=1e-12, atol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GRIDS, ids=os.path.basename)
def test_k3_matches_golden(path):
    from bpl_next_b200 import score_grid_host

    z = np.load(path)
    model, s, fx, mg = _grid_inputs(z)
    grid, outcome = score_grid_host(model, s, fx, mg)
    np.testing.assert_allclose(grid, z["grid"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(outcome, z["outcome"], rtol=0, atol=5e-6)
