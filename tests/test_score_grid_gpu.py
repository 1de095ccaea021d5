"""GPU parity: K3 (score grid / outcome) vs the numpy restatement of the reference predict path.

Tolerance is BASELINE.json's: predictive grids within 1e-6 absolute."""
import numpy as np
import pytest

from oracle import datasets, predict as op

pytestmark = pytest.mark.gpu
ATOL = 1e-6


def make_samples(model, S, T, Cf, seed, spread=0.3):
    rng = np.random.default_rng(seed)
    f32 = np.float32
    s = {"attack": rng.normal(0, spread, (S, T)).astype(f32), "defence": rng.normal(0, spread, (S, T)).astype(f32),
         "corr_coef": rng.uniform(-0.15, 0.15, S).astype(f32)}
    if model == "dixon_coles":
        s["home_advantage"] = rng.normal(0.25, 0.1, S).astype(f32)
    elif model == "extended":
        s["home_advantage"] = rng.normal(0.25, 0.1, (S, T)).astype(f32)
    else:
        for k, m in (("home_attack", 0.1), ("away_attack", -0.1), ("home_defence", 0.1), ("away_defence", -0.1)):
            s[k] = rng.normal(m, 0.1, (S, T)).astype(f32)
        if model == "neutral_wc":
            s["confederation_strength"] = rng.normal(0, 0.2, (S, Cf)).astype(f32)
    return s


def make_fixtures(model, F, T, Cf, seed):
    rng = np.random.default_rng(seed)
    h = rng.integers(0, T, F)
    a = (h + rng.integers(1, T, F)) % T
    fx = {"home_team": h.astype(np.uint16), "away_team": a.astype(np.uint16)}
    if model in ("neutral", "neutral_wc"):
        fx["neutral_venue"] = (rng.random(F) < 0.4).astype(np.uint8)
    if model == "neutral_wc":
        fx["home_conf"] = rng.integers(0, Cf, F).astype(np.uint8)
        fx["away_conf"] = rng.integers(0, Cf, F).astype(np.uint8)
    return fx


def oracle_grid(model, s, fx, max_goals):
    kw = {k: fx[k] for k in ("home_conf", "away_conf", "neutral_venue") if k in fx}
    grid, HG, AG = op.predict_score_grid_proba(model, s, fx["home_team"], fx["away_team"], max_goals, **kw)
    out = op.predict_outcome_proba(model, s, fx["home_team"], fx["away_team"], max_goals, **kw)
    return grid, np.stack([out["home_win"], out["draw"], out["away_win"]], axis=1)


@pytest.mark.parametrize("model", ["dixon_coles", "extended", "neutral", "neutral_wc"])
@pytest.mark.parametrize("max_goals,S,T,F", [(10, 203, 7, 37), (15, 64, 21, 300), (5, 9, 3, 5), (20, 40, 5, 11)])
def test_grid_device(model, max_goals, S, T, F):
    import torch
    from bpl_next_b200 import score_grid

    Cf = 4
    s = make_samples(model, S, T, Cf, seed=1)
    fx = make_fixtures(model, F, T, Cf, seed=2)
    ds = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
    dfx = {k: torch.from_numpy(v).cuda() for k, v in fx.items()}
    grid, outcome = score_grid(model, ds, dfx, max_goals)
    torch.cuda.synchronize()
    g_o, o_o = oracle_grid(model, s, fx, max_goals)
    np.testing.assert_allclose(grid.cpu().numpy(), g_o, rtol=0, atol=ATOL)
    np.testing.assert_allclose(outcome.cpu().numpy(), o_o, rtol=0, atol=5e-6)


def test_grid_host_and_reference_properties():
    """Host entry point + the assertions of the reference's tests/test_base_models.py:33-47."""
    from bpl_next_b200 import score_grid_host

    model, S, T, F = "extended", 500, 20, 40
    s = make_samples(model, S, T, 0, seed=7)
    fx = make_fixtures(model, F, T, 0, seed=8)
    grid, outcome = score_grid_host(model, s, fx, 15)
    assert np.all(grid >= 0) and np.all(grid <= 1)
    np.testing.assert_allclose(outcome.sum(axis=1), 1.0, atol=1e-5)
    g_o, o_o = oracle_grid(model, s, fx, 15)
    np.testing.assert_allclose(grid, g_o, rtol=0, atol=ATOL)


def test_grid_negative_tau_is_clipped():
    """corr_coef outside the fitted bounds: tau < 0 must give probability 0 (bpl/_util.py:62-68)."""
    from bpl_next_b200 import score_grid_host

    model, S, T, F = "dixon_coles", 16, 4, 6
    s = make_samples(model, S, T, 0, seed=3)
    s["corr_coef"] = np.full(S, 5.0, dtype=np.float32)  # 1 - 5 * lh * la < 0
    fx = make_fixtures(model, F, T, 0, seed=4)
    grid, _ = score_grid_host(model, s, fx, 6)
    g_o, _ = oracle_grid(model, s, fx, 6)
    assert np.all(grid[:, 0, 0] == 0.0)
    np.testing.assert_allclose(grid, g_o, rtol=0, atol=ATOL)


def test_grid_config5_sample_sharding_is_additive():
    """BASELINE config 5 shape (reduced S, F): two sample shards with scale 1/S_total sum to the full grid."""
    import torch
    from bpl_next_b200 import score_grid

    s, fx = datasets.config_5(S=1024, F=600)
    ds = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
    dfx = {k: torch.from_numpy(v).cuda() for k, v in fx.items()}
    full, out_full = score_grid("neutral_wc", ds, dfx, 10)
    half = [{k: v[i * 512:(i + 1) * 512].contiguous() for k, v in ds.items()} for i in range(2)]
    parts = [score_grid("neutral_wc", h, dfx, 10, scale=1.0 / 1024) for h in half]
    torch.cuda.synchronize()
    np.testing.assert_allclose((parts[0][0] + parts[1][0]).cpu().numpy(), full.cpu().numpy(), atol=2e-7)
    np.testing.assert_allclose((parts[0][1] + parts[1][1]).cpu().numpy(), out_full.cpu().numpy(), atol=2e-6)
    sub = slice(0, 40)
    g_o, o_o = oracle_grid("neutral_wc", s, {k: v[sub] for k, v in fx.items()}, 10)
    np.testing.assert_allclose(full[sub].cpu().numpy(), g_o, rtol=0, atol=ATOL)


def test_grid_fixture_ranges_reuse_the_tables():
    """ShardedScoreGrid cuts the fixtures into ranges (to overlap the all-reduce of one range with the next range's
    kernel); ranges after the first reuse the exponential tables in the workspace (BPLX_GRID_REUSE_TABLES): same grid."""
    import torch
    from bpl_next_b200 import parallel, score_grid

    s, fx = datasets.config_5(S=640, F=3000)
    ds = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
    dfx = {k: torch.from_numpy(v).cuda() for k, v in fx.items()}
    full, out_full = score_grid("neutral_wc", ds, dfx, 10)
    sg = parallel.ShardedScoreGrid("neutral_wc", ds, dfx, 10, num_samples_total=640, chunks=3)
    grid, out = sg.run()
    torch.cuda.synchronize()
    assert len(sg.ranges) == 3
    assert torch.equal(grid, full) and torch.equal(out, out_full)
    t = sg.run(timed=True)
    assert len(t) == 3 and t[0] > 0


@pytest.mark.parametrize("model", ["extended", "neutral_wc"])
def test_grid_max_goals_63_with_large_rates_stays_finite(model):
    """The largest grid the API accepts (max_goals = 63) at rates of 4-10 goals: the normalised pmf recurrence of the
    tiled kernel must neither overflow (round 1's un-normalised powers did beyond 16 goals -- ADVICE r1) nor lose the tiny
    cells' mass: every cell finite, the oracle's values to 1e-6 absolute, rows summing to the oracle's sums."""
    import torch
    from bpl_next_b200 import score_grid

    S, T, F, Cf, mg = 48, 6, 21, 3, 63
    s = make_samples(model, S, T, Cf, seed=5)
    s["attack"] = (s["attack"] + 1.7).astype(np.float32)  # rates e^1.7 = 5.5 times larger
    fx = make_fixtures(model, F, T, Cf, seed=6)
    ds = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
    dfx = {k: torch.from_numpy(v).cuda() for k, v in fx.items()}
    grid, outcome = score_grid(model, ds, dfx, mg)
    torch.cuda.synchronize()
    g, o = grid.cpu().numpy(), outcome.cpu().numpy()
    assert np.isfinite(g).all() and np.isfinite(o).all()
    g_o, o_o = oracle_grid(model, s, fx, mg)
    assert g_o.sum(axis=(1, 2)).min() > 0.99  # (the rates are large enough to matter, small enough for a 64 x 64 grid)
    np.testing.assert_allclose(g, g_o, rtol=0, atol=1e-6)
    np.testing.assert_allclose(g.sum(axis=(1, 2)), g_o.sum(axis=(1, 2)), atol=2e-5)
    np.testing.assert_allclose(o, o_o, rtol=0, atol=2e-5)
