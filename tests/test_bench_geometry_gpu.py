"""GPU parity at the launch geometry bench.py times: thousands of chains, chain-minor layout,
one CTA per group of 32 chains over many waves -- not the few-chain cluster plans the small tests exercise.

The chains checked against the float64 oracle are the committed reference-source positions (tests/golden/
ref_shim_config_{2,3}.npz, produced by running /root/reference's own `_model` under oracle/ref_shim.py) planted among
random U(-2, 2) chains, plus 64 random chains of the batch.  Tolerances: BASELINE.json (lp 1e-5, gradient 1e-4 relative)."""
import os

import numpy as np
import pytest

from oracle import datasets, models as om, predict as op
from tests import helpers as H

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _run_batch(arr, C, planted, seed, radius=2.0):
    """Evaluate C chains chain-minor; `planted` [n, D] positions go to slots spread over the batch.  Returns the
    indices to check (planted + 64 random + first / last) and host copies."""
    import torch
    from bpl_next_b200 import Problem

    p = Problem(arr)
    rng = np.random.default_rng(seed)
    theta = rng.uniform(-radius, radius, (C, p.D)).astype(np.float32)
    slots = np.linspace(0, C - 1, len(planted)).astype(np.int64) if len(planted) else np.zeros(0, np.int64)
    for s, pos in zip(slots, planted):
        theta[s] = pos.astype(np.float32)
    t = torch.from_numpy(np.ascontiguousarray(theta.T)).cuda()
    lp, grad, cc = p.logdensity(t, chain_minor=True)
    torch.cuda.synchronize()
    st = p.stats()
    idx = np.unique(np.concatenate([slots, rng.integers(0, C, 64), [0, C - 1]]))
    out = (theta[idx], lp.cpu().numpy()[idx], grad.cpu().numpy().T[idx], cc.cpu().numpy()[idx])
    all_finite = bool(torch.isfinite(lp).all().item()) and bool(torch.isfinite(grad).all().item())
    p.close()
    return slots, idx, out, st, all_finite


def _check(arr, theta, lp, grad, cc):
    lp_o, g_o, cc_o = om.log_density_and_grad(H.to_oracle(arr), theta.astype(np.float64))
    np.testing.assert_allclose(lp, lp_o, rtol=1e-5)
    np.testing.assert_allclose(cc, cc_o, rtol=1e-4, atol=1e-6)
    err = np.abs(grad - g_o) / np.abs(g_o).max(axis=1, keepdims=True)
    assert err.max() < 1e-4, err.max()


def test_config_1_at_4096_chains():
    """configs[1]: Extended, M = 1,900, K = 3, 4,096 chains chain-minor (128 CTAs, split 1)."""
    z = np.load(os.path.join(GOLDEN, "ref_shim_config_2.npz"))
    arr = H.from_training_data("extended", datasets.config_2(), epsilon=0.01)
    slots, idx, (th, lp, g, cc), st, fin = _run_batch(arr, 4096, z["theta"], seed=41)
    assert fin
    _check(arr, th, lp, g, cc)
    # the planted reference-source positions against the reference's own numbers
    pos = np.searchsorted(idx, slots)
    np.testing.assert_allclose(lp[pos], z["lp"], rtol=1e-5)
    assert (np.abs(g[pos] - z["grad"]) / np.abs(z["grad"]).max(axis=1, keepdims=True)).max() < 1e-4


def test_config_2_at_one_full_wave_and_more():
    """configs[2]: NeutralWC, M = 40,000, T = 220; 4,768 chains = 149 groups > one wave of 148 SMs: 148 CTAs, then the
    last group as a cluster (tail split)."""
    z = np.load(os.path.join(GOLDEN, "ref_shim_config_3.npz"))
    arr = H.from_training_data("neutral_wc", datasets.config_3(), epsilon=0.1)
    slots, idx, (th, lp, g, cc), st, fin = _run_batch(arr, 4768, z["theta"], seed=42)
    assert fin
    _check(arr, th, lp, g, cc)
    pos = np.searchsorted(idx, slots)
    np.testing.assert_allclose(lp[pos], z["lp"], rtol=1e-5)
    assert (np.abs(g[pos] - z["grad"]) / np.abs(z["grad"]).max(axis=1, keepdims=True)).max() < 1e-4


def test_config_2_strong_scaling_shards():
    """The per-GPU chain counts of the 32,768-chain job at 8 and 4 GPUs (4,096 / 8,192): the launch plans bench.py
    --gpus N times (incl. the cluster-split launch of the last partial wave)."""
    arr = H.from_training_data("neutral_wc", datasets.config_3(), epsilon=0.1)
    for C in (4096, 8192):
        _, idx, (th, lp, g, cc), st, fin = _run_batch(arr, C, [], seed=43 + C)
        assert fin
        _check(arr, th, lp, g, cc)


def test_config_3_at_8192_chains():
    """configs[3]: Dynamic, M = 11,400, G = 30, 8,192 chains (theta radius 0.5: see bench.py)."""
    arr = H.from_training_data("dynamic", datasets.config_4())
    _, idx, (th, lp, g, cc), st, fin = _run_batch(arr, 8192, [], seed=44, radius=0.5)
    assert fin
    _check(arr, th, lp, g, cc)


def _grid_case(max_goals, S, F, nfix=64):
    import torch
    from bpl_next_b200 import score_grid

    s, fx = datasets.config_5(S=S, F=F)
    ds = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
    dfx = {k: torch.from_numpy(v).cuda() for k, v in fx.items()}
    grid, outcome = score_grid("neutral_wc", ds, dfx, max_goals)
    torch.cuda.synchronize()
    idx = np.sort(np.random.default_rng(9).choice(F, nfix, replace=False))
    g_ref, _, _ = op.predict_score_grid_proba("neutral_wc", s, fx["home_team"][idx], fx["away_team"][idx], max_goals,
                                              home_conf=fx["home_conf"][idx], away_conf=fx["away_conf"][idx],
                                              neutral_venue=fx["neutral_venue"][idx])
    got = grid[torch.from_numpy(idx).cuda()].cpu().numpy()
    np.testing.assert_allclose(got, g_ref, rtol=0, atol=1e-6)
    np.testing.assert_allclose(outcome.sum(dim=1).cpu().numpy(), grid.sum(dim=(1, 2)).cpu().numpy(), atol=2e-6)
    assert bool(torch.isfinite(grid).all().item())
    return grid


def test_grid_config_4_full_size():
    """configs[4] at full size: S = 16,384, F = 10,000, 11 x 11; 64 random fixtures against the oracle (1e-6 abs)."""
    _grid_case(10, 16384, 10000)


def test_grid_config_4_default_max_goals():
    """The reference's default max_goals = 15 (bpl/base.py:15,78) at F = 10,000 (S reduced to keep the oracle fast)."""
    _grid_case(15, 2048, 10000, nfix=32)


def test_tail_split_matches_the_single_launch(bplx_env):
    """216 groups of configs[2] chains on 148 SMs: one full wave of single CTAs, then 68 groups as 2-CTA clusters in a
    second launch.  Same numbers as the one-launch plan (BPLX_NO_TAIL_SPLIT=1), to the rounding of the different
    cross-warp summation order; tail chains against the oracle."""
    import torch
    from bpl_next_b200 import Problem, _abi

    arr = H.from_training_data("neutral_wc", datasets.config_3(), epsilon=0.1)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    C = (sms + sms // 2 - 6) * 32 - 5  # (a ragged last group as well)
    rng = np.random.default_rng(77)
    outs = []
    for env in ({"BPLX_NO_TAIL_SPLIT": None}, {"BPLX_NO_TAIL_SPLIT": "1"}):
        bplx_env(**env)
        p = Problem(arr)
        theta = rng.uniform(-2, 2, (C, p.D)).astype(np.float32) if not outs else theta
        t = torch.from_numpy(np.ascontiguousarray(theta.T)).cuda()
        n0 = _abi.lib().bplx_launch_count()
        lp, grad, cc = p.logdensity(t, chain_minor=True)
        torch.cuda.synchronize()
        outs.append((lp.cpu().numpy(), grad.cpu().numpy().T, cc.cpu().numpy(), _abi.lib().bplx_launch_count() - n0))
        p.close()
    assert outs[0][3] == 2 and outs[1][3] == 1  # two launches with the tail split, one without
    np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=2e-6)
    np.testing.assert_allclose(outs[0][2], outs[1][2], rtol=1e-5, atol=1e-7)
    scale = np.abs(outs[1][1]).max(axis=1, keepdims=True)
    assert (np.abs(outs[0][1] - outs[1][1]) / scale).max() < 1e-5
    assert np.array_equal(outs[0][0][: sms * 32], outs[1][0][: sms * 32])  # the full wave is the same launch
    idx = np.array([sms * 32, sms * 32 + 31, C - 40, C - 1])
    _check(arr, theta[idx], outs[0][0][idx], outs[0][1][idx], outs[0][2][idx])


def test_host_entry_point_through_the_native_layout(bplx_env):
    """Large batches through bplx_logdensity_fwdbwd_host are transposed on the device to the kernel's [D, chains] layout
    (and chunked so that copies and kernels overlap): same bits as the kernel run on the [chains, D] buffers directly,
    in one piece and in two chunks, ragged chain counts included; the default chunking of a 22 MB batch and eight small
    chunks (each few enough chains for the cluster plans, which sum in another order) to rounding; a sample against the
    oracle."""
    from bpl_next_b200 import Problem

    arr = H.from_training_data("neutral_wc", datasets.config_3(), epsilon=0.1)
    rng = np.random.default_rng(91)
    for C, chunks, exact in ((4101, 1, True), (8200, 2, True), (4101, None, False), (4101, 8, False)):
        p = Problem(arr)
        theta = rng.uniform(-2, 2, (C, p.D)).astype(np.float32)
        outs = []
        for env in ({"BPLX_NO_TRANSPOSE": None, "BPLX_HOST_CHUNKS": chunks},
                    {"BPLX_NO_TRANSPOSE": "1", "BPLX_HOST_CHUNKS": 1}):
            bplx_env(**env)
            outs.append([x.copy() for x in p.logdensity_host(theta)])
        if exact:
            for a, b in zip(outs[0], outs[1]):
                assert np.array_equal(a, b)
        else:
            np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=2e-6)
            np.testing.assert_allclose(outs[0][2], outs[1][2], rtol=1e-5, atol=1e-7)
            assert (np.abs(outs[0][1] - outs[1][1]) / np.abs(outs[1][1]).max(axis=1, keepdims=True)).max() < 1e-5
        idx = np.array([0, 31, 32, C // 2, C - 1])
        _check(arr, theta[idx], outs[0][0][idx], outs[0][1][idx], outs[0][2][idx])
        p.close()


def test_chain_major_device_call_through_the_native_layout(bplx_env):
    """A large [chains, D] batch on the device (what a vmapped jax.ffi call hands over) is transposed into the kernel's
    native layout when the workspace has the room bplx_logdensity_workspace_bytes asks for: same bits as the chain-minor
    call and as the untransposed chain-major call; the likelihood-only entry point takes the same route."""
    import torch
    from bpl_next_b200 import Problem, _abi

    arr = H.from_training_data("neutral_wc", datasets.config_3(), epsilon=0.1)
    p = Problem(arr)
    C = 4101
    theta = np.random.default_rng(92).uniform(-2, 2, (C, p.D)).astype(np.float32)
    tM = torch.from_numpy(theta).cuda()
    n0 = _abi.lib().bplx_launch_count()
    a = [x.clone() for x in p.logdensity(tM)]
    assert _abi.lib().bplx_launch_count() - n0 == 3  # transpose, kernel, transpose
    b = [x.clone() for x in p.logdensity(tM.t().contiguous(), chain_minor=True)]
    bplx_env(BPLX_NO_TRANSPOSE=1)
    n0 = _abi.lib().bplx_launch_count()
    c = [x.clone() for x in p.logdensity(tM)]
    assert _abi.lib().bplx_launch_count() - n0 == 1
    torch.cuda.synchronize()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1].t()) and torch.equal(a[2], b[2])
    assert torch.equal(a[0], c[0]) and torch.equal(a[1], c[1]) and torch.equal(a[2], c[2])
    bplx_env(BPLX_NO_TRANSPOSE=None)
    Dl = int(_abi.lib().bplx_loglik_num_inputs(p._h))
    tab = torch.from_numpy(np.random.default_rng(93).uniform(0.2, 0.9, (C, Dl)).astype(np.float32)).cuda()
    n0 = _abi.lib().bplx_launch_count()
    la = [x.clone() for x in p.loglik(tab)]
    assert _abi.lib().bplx_launch_count() - n0 == 3
    lb = [x.clone() for x in p.loglik(tab.t().contiguous(), chain_minor=True)]
    torch.cuda.synchronize()
    assert torch.equal(la[0], lb[0]) and torch.equal(la[1], lb[1].t()) and torch.equal(la[2], lb[2])
    p.close()


def test_dynamic_model_chain_major_call_through_the_native_layout(bplx_env):
    """The dynamic model takes the same route: a [chains, D] batch above the transposing threshold gives the bits of the
    [D, chains] call, and of the untransposed call; a few chains against the oracle."""
    import torch
    from bpl_next_b200 import Problem, _abi

    arr = H.small_problem("dynamic", seed=4, Cf=12, T=9, M=700, neutral_frac=0.2)
    p = Problem(arr)
    C = (1 << 17) // p.D + 37
    theta = np.random.default_rng(94).uniform(-0.4, 0.4, (C, p.D)).astype(np.float32)
    tM = torch.from_numpy(theta).cuda()
    n0 = _abi.lib().bplx_launch_count()
    a = [x.clone() for x in p.logdensity(tM)]
    assert _abi.lib().bplx_launch_count() - n0 == 3
    b = [x.clone() for x in p.logdensity(tM.t().contiguous(), chain_minor=True)]
    bplx_env(BPLX_NO_TRANSPOSE=1)
    c = [x.clone() for x in p.logdensity(tM)]
    torch.cuda.synchronize()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1].t()) and torch.equal(a[2], b[2])
    assert torch.equal(a[0], c[0]) and torch.equal(a[1], c[1]) and torch.equal(a[2], c[2])
    idx = np.array([0, 31, 32, C - 1])
    _check(arr, theta[idx], a[0].cpu().numpy()[idx], a[1].cpu().numpy()[idx], a[2].cpu().numpy()[idx])
    p.close()
