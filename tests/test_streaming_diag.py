"""Streaming diagnostics (no stored draws): the finalisation formulas on CPU, the step kernel's accumulators on GPU."""
import numpy as np
import pytest
import torch

from bpl_next_b200 import diagnostics as dg


def _ar1(N, C, seed=0):
    g = torch.Generator().manual_seed(seed)
    phi = torch.tensor([0.0, 0.3, 0.6, 0.8, -0.4])[:, None]
    e = torch.randn((N, 5, C), generator=g)
    x = torch.zeros(N, 5, C)
    x[0] = e[0]
    for k in range(1, N):
        x[k] = phi * x[k - 1] + e[k] * torch.sqrt(1 - phi ** 2)
    return (x * torch.tensor([1.0, 10, 0.1, 3, 2])[:, None] + torch.tensor([100.0, -5, 0, 1e3, 0.5])[:, None]).float()


@pytest.mark.parametrize("N", [400, 401])
def test_streaming_summary_matches_stored_draw_diagnostics(N):
    x = _ar1(N, 64)
    ess, rhat = dg.effective_sample_size(x), dg.split_rhat(x)
    s = dg.streaming_summary(dg.accumulate_reference(x, 40))
    assert s["lag_window_hit"] == 0
    np.testing.assert_allclose(s["ess"].numpy(), ess.numpy(), rtol=1e-5)
    np.testing.assert_allclose(s["rhat"].numpy(), rhat.numpy(), rtol=0, atol=1e-7)
    np.testing.assert_allclose(s["mean"].numpy(), x.double().mean((0, 2)).numpy(), rtol=1e-6, atol=1e-7)
    short = dg.streaming_summary(dg.accumulate_reference(x, 6))  # window too short for phi = 0.8: flagged, never silent
    assert short["lag_window_hit"] >= 1 and short["ess"][3] > ess[3]


@pytest.mark.gpu
def test_step_kernel_accumulators_match_their_specification():
    """A run with thin = 1 stores every draw: the kernel's streaming accumulators must equal accumulate_reference of
    the stored draws (float32 sums in draw order), and the streaming ESS / R-hat the stored-draw diagnostics."""
    from bpl_next_b200 import nuts as bn

    D, C = 37, 96
    g = torch.Generator(device="cuda").manual_seed(5)
    mu = torch.randn((D, 1), generator=g, device="cuda") * 3
    sd = torch.exp(torch.randn((D, 1), generator=g, device="cuda"))

    def potential(theta, lp, grad):
        z = (theta - mu) / sd
        lp.copy_(-0.5 * (z * z).sum(0))
        grad.copy_(-z / sd)

    theta0 = torch.rand((D, C), generator=g, device="cuda") * 4 - 2
    for generic in (False, True):
        from bpl_next_b200 import _abi
        import os
        if generic:
            os.environ["BPLX_NUTS_GENERIC"] = "1"
        _abi.lib().bplx_reload_env()
        try:
            run = bn.sample(potential, theta0.clone(), num_warmup=150, num_samples=101, seed=3, diag_lags=16)
        finally:
            os.environ.pop("BPLX_NUTS_GENERIC", None)
            _abi.lib().bplx_reload_env()
        ref = dg.accumulate_reference(run.samples, 16)
        for k in ("ref", "sums", "lag", "ring", "head"):  # float32 sums in draw order vs torch's pairwise sums
            want = ref[k].cpu().numpy()
            np.testing.assert_allclose(run.diag[k].cpu().numpy(), want, rtol=2e-5, atol=2e-6 * max(1.0, np.abs(want).max()), err_msg=k)
        s = dg.streaming_summary(run.diag)
        np.testing.assert_allclose(s["rhat"].cpu().numpy(), dg.split_rhat(run.samples).cpu().numpy(), atol=1e-5)
        if s["lag_window_hit"] == 0:
            np.testing.assert_allclose(s["ess"].cpu().numpy(), dg.effective_sample_size(run.samples).cpu().numpy(), rtol=1e-3)


@pytest.mark.gpu
def test_fit_streaming_sets_reference_attributes():
    """fit_streaming: thinned draws only, diagnostics from the accumulators; the fitted attributes have the reference's
    [S, T] layout (neutral_dixon_coles_WC.py:308-334) and predict works."""
    from bpl_next_b200 import NeutralDixonColesMatchPredictorWC
    from oracle import datasets

    td = datasets.neutral_dummy_data()
    m = NeutralDixonColesMatchPredictorWC()
    out = m.fit_streaming(td, epsilon=0.2, num_warmup=200, num_samples=60, num_chains=64, thin=10, diag_lags=16,
                          set_posterior=True)
    assert out["complete"] and out["chains"] == 64 and out["stored_draws_per_chain"] == 6
    assert m.attack.shape == (64 * 6, 20) and m.confederation_strength.shape[0] == 64 * 6
    assert out["rhat_max"] < 1.2 and out["ess_min"] > 100
    p = m.predict_outcome_proba(["0", "1"], ["2", "3"], ["0", "0"], ["0", "0"], [0, 1])
    tot = p["home_win"] + p["draw"] + p["away_win"]
    assert np.all(np.abs(tot - 1) < 5e-2)
