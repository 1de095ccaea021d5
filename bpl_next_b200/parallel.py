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

    On GPUs the exchange is NOT an NCCL call: the ranks keep their partial grids in symmetric memory
    (``torch.distributed._symmetric_memory``: every rank's buffer mapped into every rank over NVLink) and one kernel per
    range and rank (``bplx_peer_sum``, ``csrc/peer_sum.cu``) exchanges flags with the peers and sums the range of every
    rank's buffer in rank order -- same bits on every rank, no staging, latency of one kernel.  The buffers are
    double-buffered by run so that a rank never overwrites a range a slower peer may still be reading.  ``peer=False``
    (or a failed rendezvous, reported in ``exchange_kind``) falls back to ``all_reduce``.
    """

    def __init__(self, model: str, local_samples: Dict[str, torch.Tensor], fixtures: Dict[str, torch.Tensor],
                 max_goals: int, num_samples_total: int, group=None, chunks: Optional[int] = None,
                 local_fn: Optional[Callable] = None, peer: bool = True, use_graph: bool = True):
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
        if F >= 16 * n:
            # the LAST range's exchange is the exposed one: it gets a third of an equal share; range starts at multiples of
            # four fixtures (16-byte aligned rows for the vector loads of the exchange)
            last = F // (3 * n) if n > 1 else F
            edges = [(F - last) * i // (n - 1) for i in range(n)] + [F] if n > 1 else [0, F]
            q = 256 if F >= 1024 * n else 4  # whole CTAs of the grid kernel (one thread per fixture, 256 per CTA)
            edges = [min(F, (e + q // 2) // q * q) for e in edges[:-1]] + [F]
            edges = sorted(set(edges))
        self.ranges = [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        self.fx = [{k: (None if v is None else v[a:b]) for k, v in fixtures.items()} for a, b in self.ranges]
        self.ws = None
        self.cuda = dev.type == "cuda"
        self.exchange_kind = "none (one rank)" if not self.dist else \
            f"all_reduce(sum) of the partial grids, {len(self.ranges)} fixture ranges, range k-1 exchanged while range k computes"
        self.peer = None
        self.graphs = None
        self.use_graph = use_graph
        if self.cuda:
            self.side = torch.cuda.Stream(device=dev, priority=-1)
            self.ev = [torch.cuda.Event() for _ in self.ranges]
            self.t = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            if self.dist and peer:
                self._setup_peer(F, g, dev)

    _FLAG_FLOATS = 64  # 256 bytes of flags in front of the data of every symmetric buffer

    def _setup_peer(self, F, g, dev):
        import ctypes as C

        from . import _abi
        try:
            import torch.distributed._symmetric_memory as symm
            grp = self.group if self.group is not None else dist.group.WORLD
            n = dist.get_world_size(grp)
            nd = (F * g * g + F * 3 + 3) // 4 * 4
            bufs = [symm.empty(self._FLAG_FLOATS + nd, dtype=torch.float32, device=dev) for _ in range(2)]
            handles = [symm.rendezvous(t, grp) for t in bufs]
            for t in bufs:
                t.zero_()
            torch.cuda.synchronize(dev)
            dist.barrier(group=grp)  # every rank's flags are zero before anyone publishes
            ptrs = [(C.c_void_p * n)(*[int(p) for p in h.buffer_ptrs]) for h in handles]
            self.peer = {
                "n": n, "rank": dist.get_rank(grp), "bufs": bufs, "handles": handles, "ptrs": ptrs, "parity": 0, "lib": _abi.lib(), "F": F, "gg": g * g,
                "grid": [t[self._FLAG_FLOATS:self._FLAG_FLOATS + F * g * g].view(F, g, g) for t in bufs],
                "outcome": [t[self._FLAG_FLOATS + F * g * g:self._FLAG_FLOATS + F * g * g + F * 3].view(F, 3) for t in bufs],
                "scratch": (torch.empty((F, g, g), dtype=torch.float32, device=dev), torch.empty((F, 3), dtype=torch.float32, device=dev)),
            }
            self.exchange_kind = (f"peer-memory sum over NVLink (bplx_peer_sum per rank and fixture range: flag exchange, then "
                                  f"every rank's partial grid read in rank order), {len(self.ranges)} fixture ranges, range k-1 "
                                  f"exchanged while range k computes")
        except Exception as e:  # no symmetric memory on this system: NCCL
            self.peer = None
            self.exchange_kind += f" [symmetric memory unavailable: {type(e).__name__}: {e}]"
        # every rank must take the same path: one that spins on peer flags next to one inside all_reduce would hang
        ok = torch.tensor([1 if self.peer is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0 and self.peer is not None:
            self.peer = None
            self.exchange_kind = (f"all_reduce(sum) of the partial grids, {len(self.ranges)} fixture ranges, range k-1 exchanged "
                                  f"while range k computes [symmetric memory unavailable on another rank]")

    def _local(self, i, exchange=True):
        a, b = self.ranges[i]
        kw = dict(scale=self.scale, want_outcome=True)
        if self.cuda:
            if self.peer is not None and not exchange:  # (the compute-only pass of a timed run: private scratch)
                kw.update(grid=self.peer["scratch"][0][a:b], outcome=self.peer["scratch"][1][a:b])
            elif self.peer is not None:  # the partial sums go to this run's symmetric buffer
                par = self.peer["parity"]
                kw.update(grid=self.peer["grid"][par][a:b], outcome=self.peer["outcome"][par][a:b])
            else:
                kw.update(grid=self.grid[a:b], outcome=self.outcome[a:b])
            kw.update(workspace=self.ws, reuse_tables=i > 0)
            self.fn(self.model, self.samples, self.fx[i], self.max_goals, **kw)
        else:  # CPU stand-in of the tests
            gr, oc = self.fn(self.model, self.samples, self.fx[i], self.max_goals, **kw)
            self.grid[a:b].copy_(gr)
            self.outcome[a:b].copy_(oc)

    def _exchange(self, i):
        a, b = self.ranges[i]
        if self.peer is not None:
            from . import _abi
            P = self.peer
            par = P["parity"]
            _abi.check(P["lib"].bplx_peer_sum(  # (epoch 0: the kernel counts the calls itself -> replayable in a graph)
                P["ptrs"][par], P["n"], P["rank"], self._FLAG_FLOATS * 4,
                a * P["gg"], (b - a) * P["gg"], self.grid[a:b].data_ptr(),
                P["F"] * P["gg"] + a * 3, (b - a) * 3, self.outcome[a:b].data_ptr(),
                0, torch.cuda.current_stream().cuda_stream))
            return
        dist.all_reduce(self.grid[a:b], op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(self.outcome[a:b], op=dist.ReduceOp.SUM, group=self.group)

    def _body(self, exchange: bool):
        """One pass over the fixture ranges on the current stream (+ the side stream for the exchanges)."""
        main = torch.cuda.current_stream()
        for i in range(len(self.ranges)):
            self._local(i, exchange)
            if self.dist and exchange:
                self.ev[i].record(main)
                self.side.wait_event(self.ev[i])
                with torch.cuda.stream(self.side):
                    self._exchange(i)
        if self.dist and exchange:
            main.wait_stream(self.side)

    def _capture(self):
        """With the peer-memory exchange a run has no argument that changes between runs: it is captured once per buffer
        parity (plus once without the exchange, for the timed split) and replayed -- every rank's ~16 launches then start
        without Python in between, which is most of the skew the flag wait would otherwise absorb."""
        self.graphs = None
        if self.peer is None or not self.use_graph:
            return
        try:
            cap = torch.cuda.Stream(device=self.grid.device)
            cap.wait_stream(torch.cuda.current_stream())
            graphs = {}
            with torch.cuda.stream(cap):
                for key, par, ex in (("x0", 0, True), ("x1", 1, True), ("c0", 0, False)):
                    self.peer["parity"] = par
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=cap):
                        self._body(ex)
                    graphs[key] = g
            torch.cuda.current_stream().wait_stream(cap)
            self.graphs = graphs
            self.peer["parity"] = 1
        except Exception as e:  # capture not possible here: plain launches
            self.graphs = None
            self.exchange_kind += f" [graph capture failed: {type(e).__name__}: {e}]"

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
            self._body(False)  # (lazy allocations inside the kernels' wrappers happen here, outside any capture)
            torch.cuda.current_stream().synchronize()
            self._capture()
        main = torch.cuda.current_stream()
        if self.peer is not None:
            self.peer["parity"] ^= 1
        if self.graphs is not None:
            g = self.graphs["x%d" % self.peer["parity"]]
            if not timed:
                g.replay()
                return self.grid, self.outcome
            # total = the replayed run; compute = the same launches without the exchange; their difference is what the
            # exchange exposes (events inside a graph cannot be timed)
            self.t[0].record(main)
            g.replay()
            self.t[1].record(main)
            self.graphs["c0"].replay()
            self.t[2].record(main)
            self.t[2].synchronize()
            total, comp = self.t[0].elapsed_time(self.t[1]), self.t[1].elapsed_time(self.t[2])
            return total, comp, max(total - comp, 0.0)
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
