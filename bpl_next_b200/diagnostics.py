"""Convergence diagnostics on stored draws: split R-hat and bulk effective sample size.

The reference never computes them (it keeps ``mcmc.get_samples()`` only, ``bpl/dixon_coles.py:118-122``); BASELINE's
ESS/s metric needs them, so they follow ``numpyro.diagnostics`` (Stan / BDA3 definitions): split-R-hat from
within/between half-chain variances; ESS from FFT autocovariances, Geyer's initial positive then initial monotone
sequence on pair sums.  Pure torch: runs on CPU (tests) and on the GPU (bench).  Draws are ``[N, ..., C]`` with the
chain axis LAST (the sampler's chain-minor layout); results have shape ``[...]``.
"""
from __future__ import annotations

import torch


def _chains_first(x: torch.Tensor) -> torch.Tensor:
    """[N, ..., C] -> [C, N, ...] float64"""
    return x.movedim(-1, 0).to(torch.float64)


def split_rhat(draws: torch.Tensor) -> torch.Tensor:
    x = _chains_first(draws)
    C, N = x.shape[:2]
    h = N // 2
    x = torch.cat([x[:, :h], x[:, N - h:]], dim=0)  # 2C half chains of length h
    var_within = x.var(dim=1, unbiased=True).mean(dim=0)
    var_between = x.mean(dim=1).var(dim=0, unbiased=True)
    var_est = var_within * (h - 1) / h + var_between
    return torch.sqrt(var_est / var_within)


def autocovariance(x: torch.Tensor) -> torch.Tensor:
    """Biased autocovariance along dim 1 of [C, N, ...] (numpyro ``autocovariance``: FFT, normalised by N)."""
    N = x.shape[1]
    M = 1 << (2 * N - 1).bit_length()
    xc = x - x.mean(dim=1, keepdim=True)
    f = torch.fft.rfft(xc, n=M, dim=1)
    ac = torch.fft.irfft(f * f.conj(), n=M, dim=1)[:, :N]
    return ac / N


def effective_sample_size(draws: torch.Tensor) -> torch.Tensor:
    x = _chains_first(draws)
    C, N = x.shape[:2]
    gamma = autocovariance(x)  # [C, N, ...]
    var_within = gamma[:, 0].mean(dim=0) * N / (N - 1.0)
    var_est = var_within * (N - 1.0) / N
    if C > 1:
        var_est = var_est + x.mean(dim=1).var(dim=0, unbiased=True)
    rho = 1.0 - (var_within - gamma.mean(dim=0)) / var_est  # [N, ...]
    rho[0] = 1.0
    K = N // 2
    pairs = rho[0:2 * K:2] + rho[1:2 * K:2]  # Geyer pair sums P_k, [K, ...]
    # initial positive sequence, then initial monotone sequence
    pairs = torch.clamp(pairs, min=0.0)
    positive = torch.cumprod((pairs > 0).to(pairs.dtype), dim=0)
    pairs = pairs * positive
    pairs = torch.cummin(pairs, dim=0).values
    tau = -1.0 + 2.0 * pairs.sum(dim=0)
    return C * N / tau


def summary(draws: torch.Tensor):
    """mean, sd, split R-hat and ESS per parameter of ``[N, D, C]`` draws."""
    x = draws.to(torch.float64)
    return {"mean": x.mean(dim=(0, -1)), "sd": x.movedim(-1, 0).reshape(-1, *x.shape[1:-1]).std(dim=0),
            "rhat": split_rhat(draws), "ess": effective_sample_size(draws)}


# ---- streaming diagnostics: no stored draws -------------------------------------------------------------------------
def accumulate_reference(draws: torch.Tensor, lags: int) -> dict:
    """What the NUTS step kernel accumulates per chain and parameter (include/bplx_nuts.h ``dg_*``), from stored draws
    ``[N, ..., C]``: the specification of the kernel's streaming accumulators, used by the tests."""
    N = draws.shape[0]
    h = N // 2
    L = max(0, min(int(lags), N - 1))
    ref = draws[0].clone()
    v = draws - ref
    sums = torch.stack([v.sum(0), (v * v).sum(0), v[:h].sum(0), (v[:h] ** 2).sum(0), v[N - h:].sum(0), (v[N - h:] ** 2).sum(0)])
    lag = torch.stack([(v[l:] * v[:N - l]).sum(0) for l in range(1, L + 1)]) if L else v[:0]
    ring = torch.zeros((L,) + tuple(draws.shape[1:]), dtype=draws.dtype, device=draws.device)
    for k in range(max(N - L, 0), N):
        ring[k % L] = v[k]
    return {"ref": ref, "sums": sums, "lag": lag, "ring": ring, "head": v[:L].clone(), "lags": L, "n": N}


def streaming_summary(diag: dict, group=None) -> dict:
    """Split R-hat, bulk ESS, posterior mean and sd per parameter from the per-chain streaming accumulators of THIS
    rank's chains; under ``torch.distributed`` the sums over chains are all-reduced (two small exchanges of
    ``(lags + 6) x D`` and ``2 x D`` doubles), so every rank returns the diagnostics of all chains of all ranks.

    Same definitions as ``split_rhat`` / ``effective_sample_size`` above (numpyro.diagnostics); the autocovariances stop
    at lag ``diag['lags']``: ``lag_window_hit`` counts the parameters whose Geyer pair sums were still positive there
    (their ESS is an upper bound -- use a larger window)."""
    import torch.distributed as dist

    N, L = int(diag["n"]), int(diag["lags"])
    h = N // 2
    f64 = torch.float64
    ref = diag["ref"].to(f64)
    S = diag["sums"].to(f64)
    C_local = ref.shape[-1]
    D = ref.shape[0]
    m_v = S[0] / N                                   # per-chain mean of the shifted draws
    mean_c = ref + m_v                               # per-chain mean
    # chain-summed autocovariances with the whole-chain mean: (1/N) [P_l - m (A_l + B_l) + (N - l) m^2],
    # A_l / B_l = the chain's sum without its first / last l values
    gam_sum = [(S[1] / N - m_v * m_v).sum(-1)]
    first = torch.zeros_like(m_v)
    last = torch.zeros_like(m_v)
    for l in range(1, L + 1):
        first = first + diag["head"][l - 1].to(f64)
        last = last + diag["ring"][(N - l) % L].to(f64)
        P = diag["lag"][l - 1].to(f64)
        gam_sum.append(((P - m_v * (2.0 * S[0] - first - last) + (N - l) * m_v * m_v) / N).sum(-1))
    gam_sum = torch.stack(gam_sum)                                          # [L+1, D]
    split = h > 1
    if split:  # half-chain statistics for split R-hat
        mh = [S[2] / h, S[4] / h]
        vh = [(S[3] - h * mh[0] ** 2) / (h - 1), (S[5] - h * mh[1] ** 2) / (h - 1)]
        half_means = torch.stack([ref + mh[0], ref + mh[1]])                # [2, D, C]
    packed = [torch.full((1, D), float(C_local), dtype=f64, device=ref.device), mean_c.sum(-1)[None], gam_sum]
    if split:
        packed += [(vh[0] + vh[1]).sum(-1)[None], half_means.sum((0, -1))[None]]
    tot = torch.cat(packed, dim=0).contiguous()
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if multi:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    C = float(tot[0, 0].item())
    grand = tot[1] / C                                                      # posterior mean
    # second exchange: between-chain sums of squares centred on the GLOBAL means (no cancellation)
    sq = [((mean_c - grand[:, None]) ** 2).sum(-1)[None]]
    if split:
        gh = tot[L + 4] / (2.0 * C)
        sq.append(((half_means - gh[None, :, None]) ** 2).sum((0, -1))[None])
    sq = torch.cat(sq, dim=0).contiguous()
    if multi:
        dist.all_reduce(sq, op=dist.ReduceOp.SUM, group=group)
    gamma_mean = tot[2:L + 3] / C                                           # [L+1, D]
    var_within = gamma_mean[0] * N / (N - 1.0)
    var_est = var_within * (N - 1.0) / N
    if C > 1:
        var_est = var_est + sq[0] / (C - 1.0)
    rho = 1.0 - (var_within[None] - gamma_mean) / var_est[None]
    rho[0] = 1.0
    K = (L + 1) // 2
    pairs = rho[0:2 * K:2] + rho[1:2 * K:2]
    pairs = torch.clamp(pairs, min=0.0)
    positive = torch.cumprod((pairs > 0).to(pairs.dtype), dim=0)
    pairs = pairs * positive
    pairs = torch.cummin(pairs, dim=0).values
    tau = -1.0 + 2.0 * pairs.sum(dim=0)
    out = {"mean": grand, "sd": torch.sqrt(gamma_mean[0] + sq[0] / C),
           "ess": C * N / tau, "num_chains": int(round(C)), "num_draws": N, "lags": L,
           "lag_window_hit": int((positive[-1] > 0).sum().item()) if K > 0 else 0}
    if split:
        W = tot[L + 3] / (2.0 * C)
        B = sq[1] / (2.0 * C - 1.0)
        out["rhat"] = torch.sqrt((W * (h - 1) / h + B) / W)
    return out
