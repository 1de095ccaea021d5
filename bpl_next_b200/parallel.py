"""One process per GPU: how the two kernels shard.

* fit: chains are independent -> contiguous chain shards, static plan replicated, no collective in the hot
  loop; only diagnostics (moments for R-hat / ESS) are all-reduced.
* predict: posterior samples are sharded; every rank computes its partial grid with ``scale = 1 / S_total`` and
  one ``all_reduce(sum)`` of ``[F, g, g]`` (+ ``[F, 3]``) finishes the mean over samples
  (the reference's ``.mean(axis=0)``, ``bpl/dixon_coles.py:163``).

``torch.distributed`` is plumbing only: NCCL on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(n: int, rank: int, nranks: int) -> Tuple[int, int]:
    """Contiguous shard (start, count) of ``n`` independent units (chains, posterior samples); the first
    ``n % nranks`` ranks take one extra."""
    if not (0 <= rank < nranks):
        raise ValueError(f"rank {rank} outside [0, {nranks})")
    base, extra = divmod(n, nranks)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def score_grid_sharded(model: str, local_samples: Dict[str, torch.Tensor], fixtures: Dict[str, torch.Tensor],
                       max_goals: int, num_samples_total: int, group=None,
                       local_fn: Optional[Callable] = None, want_outcome: bool = True):
    """Grid / outcome over sample shards.  ``local_samples`` is this rank's ``[S_local, T]`` slice.
    ``local_fn`` defaults to the CUDA kernel (``bpl_next_b200.score_grid``); tests inject a CPU stand-in."""
    if local_fn is None:
        from .problem import score_grid as local_fn
    grid, outcome = local_fn(model, local_samples, fixtures, max_goals, scale=1.0 / float(num_samples_total),
                             want_outcome=want_outcome)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grid, op=dist.ReduceOp.SUM, group=group)
        if outcome is not None:
            dist.all_reduce(outcome, op=dist.ReduceOp.SUM, group=group)
    return grid, outcome


def allreduce_chain_moments(sum_x: torch.Tensor, sum_x2: torch.Tensor, sum_mean2: torch.Tensor, num_chains: int,
                            group=None):
    """Sums of per-chain statistics over all ranks (inputs are this rank's sums over its chains, ``[D]`` each):
    what split-R-hat needs from the other GPUs.  Returns the global sums and chain count."""
    n = torch.tensor([float(num_chains)], dtype=torch.float64, device=sum_x.device)
    packed = torch.cat([sum_x.double().flatten(), sum_x2.double().flatten(), sum_mean2.double().flatten(), n])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    D = sum_x.numel()
    return packed[:D], packed[D:2 * D], packed[2 * D:3 * D], int(round(packed[-1].item()))
