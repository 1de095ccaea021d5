"""CPU: the static plan (product host code) evaluated in double == the oracle restatement."""
import numpy as np
import pytest

from oracle import datasets, models as om
from tests import helpers as H

CASES = [
    ("dixon_coles", dict()),
    ("extended", dict(weighted=False)),
    ("extended", dict(weighted=True, K=3)),
    ("neutral", dict(K=2)),
    ("neutral", dict(neutral_frac=0.0)),
    ("neutral", dict(neutral_frac=1.0)),
    ("neutral_wc", dict()),
    ("neutral_wc", dict(multi_conf=True, K=2)),
    ("dynamic", dict(Cf=4, T=5, M=90)),
    ("dynamic", dict(Cf=3, T=6, M=120, K=2, neutral_frac=0.0)),
    ("dynamic", dict(Cf=5, T=4, M=40, as_written=True)),
]


@pytest.mark.parametrize("model,kw", CASES)
@pytest.mark.parametrize("radius", [0.5, 2.0])
def test_plan_matches_oracle(model, kw, radius):
    kw = dict(kw)
    as_written = kw.pop("as_written", False)
    arr = H.small_problem(model, seed=3, **kw)
    arr.as_written = as_written
    d = H.to_oracle(arr)
    D = om.num_params(model, arr.num_teams, arr.num_covariates, arr.num_conferences, arr.num_gameweeks)
    theta = H.random_theta(D, 6, seed=11, radius=radius)
    lp_o, g_o, cc_o = om.log_density_and_grad(d, theta)
    lp_p, g_p, cc_p = H.plancheck_eval(arr, theta)
    # float32 weights/covariates are shared; everything else is double on both sides
    np.testing.assert_allclose(lp_p, lp_o, rtol=1e-7)  # const_term, yexp are float32
    np.testing.assert_allclose(cc_p, cc_o, rtol=1e-9, atol=1e-12)
    scale = np.abs(g_o).max(axis=1, keepdims=True)
    np.testing.assert_allclose(g_p / scale, g_o / scale, rtol=0, atol=2e-7)


def test_plan_reference_fixtures():
    for model, td, eps in [("dixon_coles", datasets.dummy_data(), None),
                           ("extended", datasets.timed_dummy_data(), 1.0),
                           ("neutral", datasets.neutral_dummy_data(), 0.5),
                           ("neutral_wc", datasets.neutral_dummy_data(), 0.2)]:
        arr = H.from_training_data(model, td, epsilon=eps)
        d = H.to_oracle(arr)
        D = om.num_params(model, arr.num_teams, arr.num_covariates, arr.num_conferences)
        theta = H.random_theta(D, 3, seed=5, radius=1.0)
        lp_o, g_o, cc_o = om.log_density_and_grad(d, theta)
        lp_p, g_p, cc_p = H.plancheck_eval(arr, theta)
        np.testing.assert_allclose(lp_p, lp_o, rtol=1e-7)
        np.testing.assert_allclose(cc_p, cc_o, rtol=1e-9, atol=1e-12)
        scale = np.abs(g_o).max(axis=1, keepdims=True)
        np.testing.assert_allclose(g_p / scale, g_o / scale, rtol=0, atol=2e-7)


def _edge_cases():
    """Ragged inputs the reference's data prep can produce: one match, a team that never plays, zero weights,
    double-digit goals, every match at a neutral venue, a single confederation."""
    import numpy as np
    from bpl_next_b200 import data as bdata
    out = []
    one = bdata.MatchArrays(model="dixon_coles", num_teams=2, home_team=np.array([1], np.uint16),
                            away_team=np.array([0], np.uint16), home_goals=np.array([0], np.uint8),
                            away_goals=np.array([0], np.uint8))
    out.append(("one_match", one))
    idle = H.small_problem("extended", seed=4, T=6, M=30, K=2)
    idle.num_teams = 9  # teams 6..8 never play
    idle.covariates = np.vstack([idle.covariates, np.zeros((3, 2), np.float32)])
    out.append(("idle_teams", idle))
    zw = H.small_problem("neutral", seed=5, T=5, M=40)
    zw.weights = zw.weights.copy()
    zw.weights[::3] = 0.0
    zw.home_goals = zw.home_goals.copy()
    zw.home_goals[1] = 12
    out.append(("zero_weights_big_score", zw))
    wc1 = H.small_problem("neutral_wc", seed=6, T=4, M=25, Cf=1, neutral_frac=1.0)
    out.append(("wc_single_conf_all_neutral", wc1))
    return out


@pytest.mark.parametrize("name,arr", _edge_cases(), ids=[n for n, _ in _edge_cases()])
def test_plan_edge_cases(name, arr):
    d = H.to_oracle(arr)
    D = om.num_params(arr.model, arr.num_teams, arr.num_covariates, arr.num_conferences, arr.num_gameweeks)
    theta = H.random_theta(D, 4, seed=13, radius=1.0)
    lp_o, g_o, cc_o = om.log_density_and_grad(d, theta)
    lp_p, g_p, cc_p = H.plancheck_eval(arr, theta)
    np.testing.assert_allclose(lp_p, lp_o, rtol=1e-7)
    np.testing.assert_allclose(cc_p, cc_o, rtol=1e-9, atol=1e-12)
    scale = np.abs(g_o).max(axis=1, keepdims=True)
    np.testing.assert_allclose(g_p / scale, g_o / scale, rtol=0, atol=2e-7)


@pytest.mark.parametrize("split_idx", [1, 2, 3])
def test_split_plans_match_oracle(split_idx):
    """The streams built for 2, 4 and 8 CTAs per chain group (few-chain cluster mode) describe the same model."""
    for model, kw in [("extended", dict(K=2)), ("neutral_wc", dict(multi_conf=True, T=13, M=400))]:
        arr = H.small_problem(model, seed=8, **kw)
        d = H.to_oracle(arr)
        D = om.num_params(model, arr.num_teams, arr.num_covariates, arr.num_conferences)
        theta = H.random_theta(D, 3, seed=19, radius=1.0)
        lp_o, g_o, cc_o = om.log_density_and_grad(d, theta)
        lp_p, g_p, cc_p = H.plancheck_eval(arr, theta, split_idx)
        np.testing.assert_allclose(lp_p, lp_o, rtol=1e-7)
        scale = np.abs(g_o).max(axis=1, keepdims=True)
        np.testing.assert_allclose(g_p / scale, g_o / scale, rtol=0, atol=2e-7)
