"""Host entry point (bplx_logdensity_fwdbwd_host): time per call by number of pipelined chunks, at the chain counts of the
strong-scaling shards of configs[2].  usage: python scripts/host_chunks_time.py [chains ...]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import bench
from bpl_next_b200 import Problem, _abi

arr, _, desc = bench.workload("cfg3")
p = Problem(arr)
for C in [int(x) for x in sys.argv[1:]] or [4096, 8192, 16384]:
    th = torch.empty((C, p.D), dtype=torch.float32).pin_memory()
    th.numpy()[:] = np.random.default_rng(1).uniform(-2, 2, (C, p.D)).astype(np.float32)
    lp = torch.empty(C, dtype=torch.float32).pin_memory()
    gr = torch.empty((C, p.D), dtype=torch.float32).pin_memory()
    cc = torch.empty(C, dtype=torch.float32).pin_memory()
    for ch in (0, 1, 2, 4, 8, 16):
        os.environ.pop("BPLX_HOST_CHUNKS", None)
        if ch:
            os.environ["BPLX_HOST_CHUNKS"] = str(ch)
        _abi.lib().bplx_reload_env()
        for _ in range(3):
            p.logdensity_host(th.numpy(), lp=lp.numpy(), grad=gr.numpy(), corr_coef=cc.numpy())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 20
        for _ in range(n):
            p.logdensity_host(th.numpy(), lp=lp.numpy(), grad=gr.numpy(), corr_coef=cc.numpy())
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print(json.dumps({"chains": C, "chunks": ch or "default", "ms_per_call": 1e3 * dt, "match_evals_per_s": C * arr.num_matches / dt,
                          "MB_each_way": C * p.D * 4 / 1e6}), flush=True)
