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
