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
DENSITY = sorted(f for f in glob.glob(os.path.join(GOLD, "*.npz"))
                 if not os.path.basename(f).startswith(("grid_", "ref_shim_", "refgrid_", "refdyn_", "posterior_")))
REF_SOURCE = sorted(glob.glob(os.path.join(GOLD, "ref_shim_*.npz")))
REF_GRIDS = sorted(glob.glob(os.path.join(GOLD, "refgrid_*.npz")))
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


# ---- vectors from the reference's own source ------------------------------------------------------------------------
# tests/golden/ref_shim_*.npz were written by scripts/make_ref_shim_golden.py in the build container: the reference's
# real fit() (its own data preparation) and _model, executed under stand-ins for the jax / numpyro names they use
# (oracle/ref_shim.py), traced at six positions of radius 0.3 ... 2.  Log-density and gradient in numpyro's site layout.
def _ref_case(path):
    from oracle import datasets

    name = os.path.basename(path)[len("ref_shim_"):-len(".npz")]
    model, td, kw = datasets.ref_shim_cases()[name]
    arr, _ = bdata.prepare(model, td, epsilon=kw.get("epsilon"), rescale_weights=kw.get("rescale_weights", False))
    return arr, np.load(path)


def test_reference_source_vectors_present():
    assert len(REF_SOURCE) == 8


@pytest.mark.parametrize("path", REF_SOURCE, ids=os.path.basename)
def test_oracle_matches_reference_source(path):
    """The oracle restatement (and the double-precision plan walker) against the reference's model + data-prep code.
    Unweighted models agree to rounding; with weights / covariates the product's host prep keeps them in float32."""
    arr, z = _ref_case(path)
    theta = z["theta"]
    lp, g, _ = om.log_density_and_grad(H.to_oracle(arr), theta)
    np.testing.assert_allclose(lp, z["lp"], rtol=2e-7)
    scale = np.abs(z["grad"]).max(axis=1, keepdims=True)
    assert (np.abs(g - z["grad"]) / scale).max() < 1e-6
    lp_p, g_p, _ = H.plancheck_eval(arr, theta)
    np.testing.assert_allclose(lp_p, z["lp"], rtol=1e-6)
    assert (np.abs(g_p - z["grad"]) / scale).max() < 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("path", REF_SOURCE, ids=os.path.basename)
def test_k1_matches_reference_source(path):
    """K1 against the reference's own source at BASELINE.json's tolerances (1e-5 / 1e-4 relative)."""
    from bpl_next_b200 import Problem

    arr, z = _ref_case(path)
    p = Problem(arr)
    lp, grad, _ = p.logdensity_host(z["theta"].astype(np.float32))
    np.testing.assert_allclose(lp, z["lp"], rtol=1e-5)
    scale = np.abs(z["grad"]).max(axis=1, keepdims=True)
    assert (np.abs(grad - z["grad"]) / scale).max() < 1e-4
    p.close()


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


# grids from the reference's predict source (scripts/make_ref_shim_grid_golden.py), same inputs as grid_<model>.npz
@pytest.mark.parametrize("path", REF_GRIDS, ids=os.path.basename)
def test_predict_oracle_matches_reference_source(path):
    z = np.load(path)
    model, s, fx, mg = _grid_inputs(z)
    kw = {k: fx[k] for k in ("neutral_venue", "home_conf", "away_conf") if k in fx}
    grid, _, _ = op.predict_score_grid_proba(model, s, fx["home_team"], fx["away_team"], mg, **kw)
    out = op.predict_outcome_proba(model, s, fx["home_team"], fx["away_team"], mg, **kw)
    np.testing.assert_allclose(grid, z["grid"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.stack([out["home_win"], out["draw"], out["away_win"]], 1), z["outcome"], rtol=0, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("path", REF_GRIDS, ids=os.path.basename)
def test_k3_matches_reference_source(path):
    """K3 against the reference's predict source at BASELINE.json's tolerance (1e-6 absolute)."""
    from bpl_next_b200 import score_grid_host

    z = np.load(path)
    model, s, fx, mg = _grid_inputs(z)
    grid, outcome = score_grid_host(model, s, fx, mg)
    np.testing.assert_allclose(grid, z["grid"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(outcome, z["outcome"], rtol=0, atol=2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("path", REF_GRIDS, ids=os.path.basename)
def test_predictor_api_matches_reference_source(path):
    """The mirrored predictor classes, given the same posterior samples as the reference's classes, against what the
    reference's own predict_score_n_proba / predict_concede_n_proba / predict_score_proba / knockout code returned."""
    import bpl_next_b200 as bp

    z = np.load(path)
    model, s, fx, mg = _grid_inputs(z)
    cls = {"dixon_coles": bp.DixonColesMatchPredictor, "extended": bp.ExtendedDixonColesMatchPredictor,
           "neutral": bp.NeutralDixonColesMatchPredictor, "neutral_wc": bp.NeutralDixonColesMatchPredictorWC}[model]
    m = cls()
    T = s["attack"].shape[1]
    m.teams = np.array([str(i) for i in range(T)])
    m._teams_dict = {str(i): i for i in range(T)}
    for k, v in s.items():
        setattr(m, k, v)
    h, a = [str(i) for i in fx["home_team"]], [str(i) for i in fx["away_team"]]
    n = np.arange(mg + 1)
    one, zero = np.ones(len(h), dtype=int), np.zeros(len(h), dtype=int)
    if model == "neutral_wc":
        Cf = s["confederation_strength"].shape[1]
        m.conferences = np.array([str(i) for i in range(Cf)])
        m._conferences_dict = {str(i): i for i in range(Cf)}
        hc, ac, nv = [str(i) for i in fx["home_conf"]], [str(i) for i in fx["away_conf"]], fx["neutral_venue"].astype(int)
        got = dict(score_home=m.predict_score_n_proba(n, h[0], a[0], hc[0], ac[0], home=True, neutral_venue=int(nv[0]), max_goals=mg),
                   score_away=m.predict_score_n_proba(n, a[0], h[0], ac[0], hc[0], home=False, neutral_venue=int(nv[0]), max_goals=mg),
                   concede_home=m.predict_concede_n_proba(n, h[0], a[0], hc[0], ac[0], home=True, neutral_venue=int(nv[0]), max_goals=mg),
                   concede_away=m.predict_concede_n_proba(n, a[0], h[0], ac[0], hc[0], home=False, neutral_venue=int(nv[0]), max_goals=mg),
                   score_1_0=m.predict_score_proba(h, a, hc, ac, one, zero, nv))
        ko = m.predict_outcome_proba(h, a, hc, ac, nv, knockout=True, max_goals=mg)
    elif model == "neutral":
        nv = fx["neutral_venue"].astype(int)
        got = dict(score_home=m.predict_score_n_proba(n, h[0], a[0], home=True, neutral_venue=int(nv[0]), max_goals=mg),
                   score_away=m.predict_score_n_proba(n, a[0], h[0], home=False, neutral_venue=int(nv[0]), max_goals=mg),
                   concede_home=m.predict_concede_n_proba(n, h[0], a[0], home=True, neutral_venue=int(nv[0]), max_goals=mg),
                   concede_away=m.predict_concede_n_proba(n, a[0], h[0], home=False, neutral_venue=int(nv[0]), max_goals=mg),
                   score_1_0=m.predict_score_proba(h, a, one, zero, nv))
        ko = m.predict_outcome_proba(h, a, nv, knockout=True, max_goals=mg)
    else:
        got = dict(score_home=m.predict_score_n_proba(n, h[0], a[0], home=True, max_goals=mg),
                   score_away=m.predict_score_n_proba(n, a[0], h[0], home=False, max_goals=mg),
                   concede_home=m.predict_concede_n_proba(n, h[0], a[0], home=True, max_goals=mg),
                   concede_away=m.predict_concede_n_proba(n, a[0], h[0], home=False, max_goals=mg),
                   score_1_0=m.predict_score_proba(h, a, one, zero))
        ko = None
    for k, v in got.items():
        np.testing.assert_allclose(np.asarray(v), z["api_" + k], rtol=0, atol=2e-6, err_msg=k)
    if ko is not None:
        np.testing.assert_allclose(np.stack([ko["home_win"], ko["away_win"]], 1), z["api_knockout"], rtol=0, atol=5e-6)


@pytest.mark.parametrize("path", REF_SOURCE, ids=os.path.basename)
def test_recorded_sites_match_reference_source(path):
    """`fit()` turns the latent draws into the sites the reference records (attack, defence, venue effects, rho,
    corr_coef, ...): the mirror's transforms + `_set_posterior` against the deterministic sites of the reference's own
    model trace at the same unconstrained positions."""
    import bpl_next_b200 as bp
    from bpl_next_b200.predictors import constrain
    from oracle import datasets

    name = os.path.basename(path)[len("ref_shim_"):-len(".npz")]
    model, td, kw = datasets.ref_shim_cases()[name]
    arr, meta = bdata.prepare(model, td, epsilon=kw.get("epsilon"), rescale_weights=kw.get("rescale_weights", False))
    z = np.load(path)
    K = 0 if arr.covariates is None else arr.covariates.shape[1]
    offs = om.layout_offsets(om.site_layout(model, arr.num_teams, K, arr.num_conferences or 0, 0))
    layout = {k: (o, int(np.prod(shape)) if shape else 1, tr) for k, (o, shape, tr) in offs.items() if k != "__D__"}
    layout = {k: v for k, v in layout.items() if v[1] > 0}
    s = constrain(z["theta"], layout)
    s["corr_coef"] = om.log_density_and_grad(H.to_oracle(arr), z["theta"])[2].astype(np.float32)
    cls = {"dixon_coles": bp.DixonColesMatchPredictor, "extended": bp.ExtendedDixonColesMatchPredictor,
           "neutral": bp.NeutralDixonColesMatchPredictor, "neutral_wc": bp.NeutralDixonColesMatchPredictorWC}[model]
    m = cls()
    m._meta, m.teams, m._teams_dict = meta, meta["teams"], meta["teams_dict"]
    m._set_posterior(arr, s)
    checked = 0
    for key in z.files:
        if not key.startswith("det_"):
            continue
        got = getattr(m, key[4:])
        np.testing.assert_allclose(np.asarray(got, dtype=np.float64), z[key], rtol=2e-5, atol=2e-6, err_msg=key)
        checked += 1
    assert checked >= 3


# ---- the dynamic model's source (a7): tests/golden/refdyn_*.npz, scripts/make_ref_shim_dynamic_golden.py -------------------
# The reference's DynamicNeutralDixonColesMatchPredictor.fit() + `_model` (bpl/dynamic_dixon_coles.py:63-296) executed
# unmodified under oracle/ref_shim.py: as fitted (num_gameweeks = max(gameweek), JAX's clamped gather made explicit) and with
# num_gameweeks = max + 1.  In both the walk never reaches the rates = the product's as-written mode.
REF_DYNAMIC = sorted(glob.glob(os.path.join(GOLD, "refdyn_*.npz")))


def _dyn_arrays(z):
    cov = z["covariates"] if "covariates" in z.files else None
    return bdata.MatchArrays(model="dynamic", num_teams=int(z["num_teams"]), home_team=z["home_team"], away_team=z["away_team"],
                             home_goals=z["home_goals"], away_goals=z["away_goals"], neutral_venue=z["neutral_venue"],
                             gameweek=z["gameweek"], num_gameweeks=int(z["num_gameweeks"]), covariates=cov, as_written=True)


def test_dynamic_reference_source_vectors_present():
    assert len(REF_DYNAMIC) == 4


@pytest.mark.parametrize("path", REF_DYNAMIC, ids=os.path.basename)
def test_oracle_matches_dynamic_reference_source(path):
    z = np.load(path)
    arr = _dyn_arrays(z)
    lp, g, _ = om.log_density_and_grad(H.to_oracle(arr), z["theta"])
    np.testing.assert_allclose(lp, z["lp"], rtol=1e-12)
    scale = np.abs(z["grad"]).max(axis=1, keepdims=True)
    assert (np.abs(g - z["grad"]) / scale).max() < 1e-12
    lp_p, g_p, _ = H.plancheck_eval(arr, z["theta"])
    np.testing.assert_allclose(lp_p, z["lp"], rtol=1e-6)
    assert (np.abs(g_p - z["grad"]) / scale).max() < 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("path", REF_DYNAMIC, ids=os.path.basename)
def test_k1d_matches_dynamic_reference_source(path):
    from bpl_next_b200 import Problem

    z = np.load(path)
    p = Problem(_dyn_arrays(z))
    lp, grad, _ = p.logdensity_host(z["theta"].astype(np.float32))
    np.testing.assert_allclose(lp, z["lp"], rtol=1e-5)
    scale = np.abs(z["grad"]).max(axis=1, keepdims=True)
    assert (np.abs(grad - z["grad"]) / scale).max() < 1e-4
    p.close()
