"""CPU, world_size 2 over gloo: the sharding + reduction logic of the N>1 paths (no GPU, no kernel)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bpl_next_b200 import parallel
from oracle import predict as op


def test_shard_partitions():
    for n in (0, 1, 7, 32768, 16384 + 3):
        for w in (1, 2, 3, 8):
            parts = [parallel.shard(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (s0, c0), (s1, _) in zip(parts, parts[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        parallel.shard(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cpu_grid(model, samples, fixtures, max_goals, scale=None, want_outcome=True):
    """CPU stand-in with the kernel's contract: scale * sum over local samples."""
    s = {k: v.numpy() for k, v in samples.items()}
    S = s["attack"].shape[0]
    grid, HG, AG = op.predict_score_grid_proba(model, s, fixtures["home_team"].numpy(), fixtures["away_team"].numpy(), max_goals)
    grid = torch.from_numpy(grid * S * scale)
    out = torch.stack([grid[:, HG > AG].sum(-1), grid[:, HG == AG].sum(-1), grid[:, HG < AG].sum(-1)], 1)
    return grid, (out if want_outcome else None)


def _worker(rank, nranks, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=nranks)
    try:
        rng = np.random.default_rng(0)  # same data on every rank
        S, T, F = 37, 5, 6
        full = {"attack": rng.normal(0, 0.3, (S, T)), "defence": rng.normal(0, 0.3, (S, T)),
                "home_advantage": rng.normal(0.2, 0.1, (S, T)), "corr_coef": rng.uniform(-0.1, 0.1, S)}
        fx = {"home_team": torch.tensor([0, 1, 2, 3, 4, 0]), "away_team": torch.tensor([1, 2, 3, 4, 0, 2])}
        start, cnt = parallel.shard(S, rank, nranks)
        local = {k: torch.from_numpy(v[start:start + cnt]) for k, v in full.items()}
        grid, out = parallel.score_grid_sharded("extended", local, fx, 10, S, local_fn=_cpu_grid)
        ref, _, _ = op.predict_score_grid_proba("extended", full, fx["home_team"].numpy(), fx["away_team"].numpy(), 10)
        ok_grid = bool(np.allclose(grid.numpy(), ref, atol=1e-12))
        # chain moments: rank r owns chains shard(r); global sums must equal the single-process sums
        C, D = 9, 4
        draws = np.random.default_rng(1).normal(size=(C, 50, D))
        c0, cn = parallel.shard(C, rank, nranks)
        mine = torch.from_numpy(draws[c0:c0 + cn])
        sx, sx2, sm2, n = parallel.allreduce_chain_moments(mine.sum((0, 1)), (mine ** 2).sum((0, 1)),
                                                          (mine.mean(1) ** 2).sum(0), cn)
        ok_mom = n == C and np.allclose(sx.numpy(), draws.sum((0, 1))) and np.allclose(sx2.numpy(), (draws ** 2).sum((0, 1))) \
            and np.allclose(sm2.numpy(), (draws.mean(1) ** 2).sum(0))
        # the overlapped variant (fixture ranges exchanged on a side stream on the GPU; sequential on CPU): same grid
        sg = parallel.ShardedScoreGrid("extended", local, fx, 10, S, chunks=3, local_fn=_cpu_grid)
        grid2, out2 = sg.run()
        ok_grid = ok_grid and len(sg.ranges) == 3 and bool(np.allclose(grid2.numpy(), ref, atol=1e-6)) \
            and bool(np.allclose(out2.numpy(), out.numpy(), atol=1e-6))
        # streaming diagnostics: per-rank accumulators, sums over chains all-reduced = the single-process diagnostics
        from bpl_next_b200 import diagnostics as dg
        g = torch.Generator().manual_seed(3)
        x = torch.cumsum(torch.randn((120, 3, 10), generator=g), 0) * 0.1 + torch.randn((120, 3, 10), generator=g)
        c0, cn = parallel.shard(10, rank, nranks)
        s_loc = dg.streaming_summary(dg.accumulate_reference(x[..., c0:c0 + cn].contiguous(), 30))
        solo = [dist.new_group([r]) for r in range(nranks)][rank]  # (every rank creates every group, in the same order)
        s_all = dg.streaming_summary(dg.accumulate_reference(x, 30), group=solo)
        ok_mom = ok_mom and s_loc["num_chains"] == 10 and np.allclose(s_loc["ess"].numpy(), s_all["ess"].numpy(), rtol=1e-9) \
            and np.allclose(s_loc["rhat"].numpy(), s_all["rhat"].numpy(), rtol=1e-12) \
            and np.allclose(s_loc["mean"].numpy(), s_all["mean"].numpy(), rtol=1e-12)
        q.put((rank, ok_grid, bool(ok_mom), float(out.sum(1).mean())))
    finally:
        dist.destroy_process_group()


def test_world_size_2_grid_allreduce_and_moments():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_grid, ok_mom, wdl in res:
        assert ok_grid and ok_mom, (rank, ok_grid, ok_mom)
        assert abs(wdl - 1.0) < 1e-3
