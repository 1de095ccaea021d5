"""Driver of the batched GPU NUTS transition (``include/bplx_nuts.h``) -- stands in for numpyro's ``NUTS`` +
``MCMC(chain_method="vectorized")`` as every reference ``fit`` builds them (``bpl/dixon_coles.py:100-116``).

    run = sample(potential, theta0, num_warmup=500, num_samples=1000)

``potential(theta_eval, lp, grad)`` evaluates log density and gradient in place on chain-minor ``[D, C]`` tensors
(for the models: ``Problem.logdensity(..., chain_minor=True)``).  For large models the sampler keeps its state
chain-major (a warp per chain, see ``csrc/nuts.cu``) and calls ``potential_cm`` on ``[C, D]`` tensors instead
(``Problem.logdensity(..., chain_minor=False)``).  torch provides device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np
import torch

from . import _abi


class Window(C.Structure):
    _fields_ = [("start", C.c_int32), ("end", C.c_int32)]


class NutsParams(C.Structure):
    _fields_ = [
        ("C", C.c_int32), ("D", C.c_int32), ("ld", C.c_int32),
        ("num_warmup", C.c_int32), ("num_samples", C.c_int32), ("thin", C.c_int32), ("num_keep", C.c_int32),
        ("max_tree_depth", C.c_int32), ("num_windows", C.c_int32),
        ("windows", C.c_void_p),
        ("target_accept", C.c_float), ("init_step_size", C.c_float), ("max_delta_energy", C.c_float),
        ("seed", C.c_uint64), ("chain_offset", C.c_int64),
        ("theta_eval", C.c_void_p), ("lp", C.c_void_p), ("grad", C.c_void_p),
        ("chain", C.c_void_p),
    ] + [(n, C.c_void_p) for n in ("p_half", "inv_mass", "zL", "rL", "gL", "zR", "rR", "gR", "zP", "gP", "r_sum", "zQ",
                                   "gQ", "r_sum_sub", "r_ckpts", "r_sum_ckpts", "wf_mean", "wf_m2", "samples",
                                   "sample_lp", "sample_accept", "active_count")] + [
        ("diag_lags", C.c_int32),
    ] + [(n, C.c_void_p) for n in ("dg_ref", "dg_sums", "dg_lag", "dg_ring", "dg_head")] + [
        ("state_layout", C.c_int32), ("ld_state", C.c_int32),
    ]


def adaptation_schedule(num_steps: int):
    """numpyro ``build_adaptation_schedule`` (Stan's windows): init buffer 75, base window 25 doubling, end buffer 50;
    15 % / 75 % / 10 % when they do not fit; a single window below 20 steps."""
    if num_steps <= 0:
        return [(0, -1)]
    if num_steps < 20:
        return [(0, num_steps - 1)]
    start_buffer, end_buffer, init_window = 75, 50, 25
    if start_buffer + end_buffer + init_window > num_steps:
        start_buffer = int(0.15 * num_steps)
        end_buffer = int(0.1 * num_steps)
        init_window = num_steps - start_buffer - end_buffer
    sched = [(0, start_buffer - 1)]
    end_window_start = num_steps - end_buffer
    next_size, next_start = init_window, start_buffer
    while next_start < end_window_start:
        cur_start, cur_size = next_start, next_size
        if 3 * cur_size <= end_window_start - cur_start:
            next_size = 2 * cur_size
        else:
            cur_size = end_window_start - cur_start
        next_start = cur_start + cur_size
        sched.append((cur_start, next_start - 1))
    sched.append((end_window_start, num_steps - 1))
    return sched


def _declare(lib):
    if getattr(lib, "_nuts_declared", False):
        return lib
    lib.bplx_nuts_chain_bytes.argtypes = []
    lib.bplx_nuts_chain_bytes.restype = C.c_size_t
    for f in (lib.bplx_nuts_init, lib.bplx_nuts_step):
        f.argtypes = [C.POINTER(NutsParams), C.c_void_p]
        f.restype = C.c_int
    lib.bplx_nuts_summary.argtypes = [C.POINTER(NutsParams), C.c_void_p]
    lib.bplx_nuts_summary.restype = C.c_int
    lib._nuts_declared = True
    return lib


@dataclass
class NutsRun:
    samples: torch.Tensor        # [num_keep, D, C] unconstrained draws (chain-minor)
    lp: torch.Tensor             # [num_keep, C]
    accept: torch.Tensor         # [num_keep, C]
    step_size: np.ndarray        # [C]
    num_divergent: np.ndarray    # [C]
    num_leapfrog: np.ndarray     # [C] over warm-up + sampling
    launches: int                # log-density evaluations = kernel launches of the potential
    inv_mass: torch.Tensor       # [D, C]
    transitions: Optional[np.ndarray] = None  # [C] transitions completed (== num_warmup + num_samples unless cut short)
    diag: Optional[dict] = None  # streaming accumulators (diagnostics.streaming_summary): ref, sums, lag, ring, head, lags, n
    block_ms: Optional[list] = None  # with time_blocks: CUDA-event time of every block of `check_every` (potential, step) pairs


def sample(potential: Callable[[torch.Tensor, torch.Tensor, torch.Tensor], None], theta0: torch.Tensor,
           num_warmup: int = 500, num_samples: int = 1000, thin: int = 1, seed: int = 42, max_tree_depth: int = 10,
           target_accept: float = 0.8, step_size: float = 1.0, chain_offset: int = 0, check_every: int = 32,
           max_launches: Optional[int] = None, use_graph: bool = True, diag_lags: int = 0,
           pad_rows: bool = False, potential_cm: Optional[Callable[[torch.Tensor, torch.Tensor, torch.Tensor], None]] = None,
           state_layout: str = "auto", time_blocks: bool = False) -> NutsRun:
    """``theta0``: ``[D, C]`` float32 CUDA tensor (chain-minor) of initial unconstrained positions.

    ``thin`` stores every thin-th post-warm-up draw only; ``diag_lags > 0`` keeps per-chain streaming accumulators of
    ALL post-warm-up draws (moments of the whole chain and of its halves, lagged products up to ``diag_lags``) inside
    the step kernel, so split R-hat and ESS need no stored draws (``diagnostics.streaming_summary``).

    ``state_layout``: ``"chain_minor"`` (every array ``[D, C]``), ``"chain_major"`` (``[C, D]``, needs ``potential_cm``)
    or ``"auto"``: chain-major whenever ``potential_cm`` is given -- measured per leapfrog of a fit: configs[0] at 1,024
    chains 29.8 -> 28.5 us, configs[1] at 4,096 chains 63 -> 40 us, configs[2] at 32,768 chains 4.57 -> 2.62 ms.  The
    returned tensors have the same logical shapes either way (views)."""
    if not theta0.is_cuda:
        raise RuntimeError("bpl_next_b200.nuts needs CUDA tensors: there is no CPU fallback")
    lib = _declare(_abi.lib())
    D, Cn = theta0.shape
    dev = theta0.device
    f32 = dict(dtype=torch.float32, device=dev)
    if state_layout not in ("auto", "chain_minor", "chain_major"):
        raise ValueError(f"state_layout {state_layout!r}")
    cm = state_layout == "chain_major" or (state_layout == "auto" and potential_cm is not None)
    if cm and potential_cm is None:
        raise ValueError("state_layout='chain_major' needs potential_cm (the log-density on [C, D] tensors)")
    # row pitch of every [D, C] array (the kernels take any pitch >= C; measured on B200: padding a 4 KB-multiple pitch
    # by one line changes nothing -- 3.03 vs 3.10 ms per step at 32,768 chains -- so the default is the dense layout)
    ld = Cn + 32 if (Cn * 4) % 4096 == 0 and pad_rows else Cn
    ldD = (D + 31) // 32 * 32  # chain-major: pitch of a chain's vector (whole 128-byte lines)

    def zeros(*lead):
        """A zeroed stack of per-(chain, parameter) vectors, as its logical ``[..., D, C]`` view."""
        if cm:
            return torch.zeros(lead + (Cn, ldD), **f32)[..., :D].transpose(-1, -2)
        return torch.zeros(lead + (D, ld), **f32)[..., :Cn]

    num_keep = (num_samples + thin - 1) // thin
    # state: 18 + 2 * max_tree_depth vectors of [D, C]; output: num_keep of them.  Fail early and say what to change.
    diag_lags = max(0, min(int(diag_lags), max(num_samples - 1, 0)))
    need = 4 * (D + 31) * (Cn + 32) * (18 + 2 * max_tree_depth + num_keep + ((7 + 3 * diag_lags) if diag_lags else 0)) + 8 * num_keep * Cn
    free, _total = torch.cuda.mem_get_info(dev)
    if need > 0.95 * free:
        raise MemoryError(
            f"NUTS on {Cn} chains x {D} parameters needs {need / 2**30:.1f} GiB ({num_keep} stored draws of "
            f"{4 * D * Cn / 2**20:.0f} MiB each), {free / 2**30:.1f} GiB are free: raise `thin` or lower num_samples / the chain count")
    vecs = {n: zeros() for n in ("p_half", "inv_mass", "zL", "rL", "gL", "zR", "rR", "gR", "zP", "gP",
                                 "r_sum", "zQ", "gQ", "r_sum_sub", "wf_mean", "wf_m2")}
    r_ckpts = zeros(max_tree_depth)
    r_sum_ckpts = zeros(max_tree_depth)
    theta_eval = zeros()
    theta_eval.copy_(theta0)
    grad = zeros()
    lp = torch.zeros(Cn, **f32)
    samples = zeros(num_keep)
    sample_lp = torch.zeros((num_keep, ld), **f32)[:, :Cn]
    sample_accept = torch.zeros((num_keep, ld), **f32)[:, :Cn]
    chain = torch.zeros(Cn * int(lib.bplx_nuts_chain_bytes()), dtype=torch.uint8, device=dev)
    active = torch.zeros(1, dtype=torch.int32, device=dev)
    sched = adaptation_schedule(num_warmup)
    windows = torch.tensor(sched, dtype=torch.int32, device=dev).contiguous()

    p = NutsParams()
    p.C, p.D, p.ld = Cn, D, ld
    p.num_warmup, p.num_samples, p.thin, p.num_keep = num_warmup, num_samples, thin, num_keep
    p.max_tree_depth, p.num_windows = max_tree_depth, len(sched)
    p.windows = windows.data_ptr()
    p.target_accept, p.init_step_size, p.max_delta_energy = target_accept, step_size, 1000.0
    p.seed, p.chain_offset = seed, chain_offset
    p.theta_eval, p.lp, p.grad = theta_eval.data_ptr(), lp.data_ptr(), grad.data_ptr()
    p.chain = chain.data_ptr()
    for n, t in vecs.items():
        setattr(p, n, t.data_ptr())
    p.r_ckpts, p.r_sum_ckpts = r_ckpts.data_ptr(), r_sum_ckpts.data_ptr()
    p.samples, p.sample_lp, p.sample_accept = samples.data_ptr(), sample_lp.data_ptr(), sample_accept.data_ptr()
    p.active_count = active.data_ptr()
    diag = None
    p.diag_lags = diag_lags
    p.state_layout, p.ld_state = (1, ldD) if cm else (0, 0)
    # what the potential sees: [D, C] views, or the chain-major [C, D] views of the same storage
    if cm:
        th_arg, gr_arg, call = theta_eval.transpose(0, 1), grad.transpose(0, 1), potential_cm
    else:
        th_arg, gr_arg, call = theta_eval, grad, potential
    if diag_lags:
        diag = {"ref": zeros(), "sums": zeros(6), "lag": zeros(diag_lags), "ring": zeros(diag_lags),
                "head": zeros(diag_lags), "lags": diag_lags, "n": num_samples}
        p.dg_ref, p.dg_sums, p.dg_lag = diag["ref"].data_ptr(), diag["sums"].data_ptr(), diag["lag"].data_ptr()
        p.dg_ring, p.dg_head = diag["ring"].data_ptr(), diag["head"].data_ptr()

    limit = max_launches if max_launches is not None else (num_warmup + num_samples + 1) * (2 ** max_tree_depth) + 8

    def block():  # check_every x (log-density, NUTS step); the last step counts the chains that are not finished
        st = torch.cuda.current_stream().cuda_stream
        for k in range(check_every):
            call(th_arg, lp, gr_arg)
            if k == check_every - 1:
                active.zero_()
            _abi.check(lib.bplx_nuts_step(C.byref(p), st))

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        _abi.check(lib.bplx_nuts_init(C.byref(p), side.cuda_stream))
        block()  # eager once: lazy allocations inside `potential` happen outside the capture
        launches = check_every
        graph = None
        if use_graph and int(active.item()) != 0:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                block()
        events = []
        while launches < limit and int(active.item()) != 0:
            if time_blocks:
                events.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
                events[-1][0].record(side)
            if graph is not None:
                graph.replay()
            else:
                block()
            if time_blocks:
                events[-1][1].record(side)
            launches += check_every
    torch.cuda.current_stream().wait_stream(side)
    summ = np.zeros((Cn, 8), dtype=np.float32)
    torch.cuda.synchronize()
    _abi.check(lib.bplx_nuts_summary(C.byref(p), summ.ctypes.data))
    return NutsRun(samples=samples, lp=sample_lp, accept=sample_accept, step_size=summ[:, 1].copy(),
                   num_divergent=summ[:, 2].astype(np.int64), num_leapfrog=summ[:, 3].astype(np.int64),
                   launches=launches, inv_mass=vecs["inv_mass"], transitions=summ[:, 0].astype(np.int64), diag=diag,
                   block_ms=[a.elapsed_time(b) for a, b in events] if time_blocks else None)
