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


class ShardedScoreGrid:
    """Predictive grid over sample shards with the exchange overlapped: the fixtures are cut into ``chunks`` ranges;
    while the kernel works on range k the ``[F_k, g, g]`` + ``[F_k, 3]`` partial sums of range k-1 are all-reduced on a
    side stream, so only the last range's exchange is exposed.  Buffers are allocated once; ``run()`` is re-entrant.

    ``run(timed=True)`` returns ``(total_ms, compute_ms, exposed_exchange_ms)`` from CUDA events on the launch stream.
    """

    def __init__(self, model: str, local_samples: Dict[str, torch.Tensor], fixtures: Dict[str, torch.Tensor],
                 max_goals: int, num_samples_total: int, group=None, chunks: Optional[int] = None,
                 local_fn: Optional[Callable] = None):
        if local_fn is None:
            from .problem import score_grid as local_fn
        self.fn, self.model, self.samples, self.group = local_fn, model, local_samples, group
        self.max_goals, self.scale = max_goals, 1.0 / float(num_samples_total)
        self.dist = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        F = int(len(fixtures["home_team"]))
        g = max_goals + 1
        dev = local_samples["attack"].device
        self.grid = torch.empty((F, g, g), dtype=torch.float32, device=dev)
        self.outcome = torch.empty((F, 3), dtype=torch.float32, device=dev)
        n = chunks if chunks is not None else (4 if self.dist and F >= 2048 else 1)
        n = max(1, min(n, F))
        edges = [F * i // n for i in range(n + 1)]
        self.ranges = [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        self.fx = [{k: (None if v is None else v[a:b]) for k, v in fixtures.items()} for a, b in self.ranges]
        self.ws = None
        self.cuda = dev.type == "cuda"
        self.exchange_kind = "none (one rank)" if not self.dist else \
            f"all_reduce(sum) of the partial grids, {len(self.ranges)} fixture ranges, range k-1 exchanged while range k computes"
        if self.cuda:
            self.side = torch.cuda.Stream(device=dev)
            self.ev = [torch.cuda.Event() for _ in self.ranges]
            self.t = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

    def _local(self, i):
        a, b = self.ranges[i]
        kw = dict(scale=self.scale, want_outcome=True)
        if self.cuda:
            kw.update(grid=self.grid[a:b], outcome=self.outcome[a:b], workspace=self.ws, reuse_tables=i > 0)
            self.fn(self.model, self.samples, self.fx[i], self.max_goals, **kw)
        else:  # CPU stand-in of the tests
            gr, oc = self.fn(self.model, self.samples, self.fx[i], self.max_goals, **kw)
            self.grid[a:b].copy_(gr)
            self.outcome[a:b].copy_(oc)

    def _exchange(self, i):
        a, b = self.ranges[i]
        dist.all_reduce(self.grid[a:b], op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(self.outcome[a:b], op=dist.ReduceOp.SUM, group=self.group)

    def run(self, timed: bool = False):
        if not self.cuda:
            for i in range(len(self.ranges)):
                self._local(i)
                if self.dist:
                    self._exchange(i)
            return self.grid, self.outcome
        if self.ws is None:  # one workspace, sized for the largest range
            from . import problem as _p
            need = max(_p.score_grid_workspace_bytes(self.model, self.samples, fx, self.max_goals) for fx in self.fx)
            self.ws = torch.empty(max(need, 1), dtype=torch.uint8, device=self.grid.device)
        main = torch.cuda.current_stream()
        if timed:
            self.t[0].record(main)
        for i in range(len(self.ranges)):
            self._local(i)
            if self.dist:
                self.ev[i].record(main)
                self.side.wait_event(self.ev[i])
                with torch.cuda.stream(self.side):
                    self._exchange(i)
        if timed:
            self.t[1].record(main)
        if self.dist:
            main.wait_stream(self.side)
        if timed:
            self.t[2].record(main)
            self.t[2].synchronize()
            total, comp = self.t[0].elapsed_time(self.t[2]), self.t[0].elapsed_time(self.t[1])
            return total, comp, max(total - comp, 0.0)
        return self.grid, self.outcome
