"""K1 per-call time, chain-minor [D, C] against chain-major [C, D] buffers (rotating through sets larger than L2).
usage: python scripts/layout_time.py [cfg3|cfg2|cfg4] [chains]"""
import json, sys
import numpy as np, torch
sys.path.insert(0, '.')
import bench
from bpl_next_b200 import Problem

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
arr, C, desc = bench.workload(wl)
if len(sys.argv) > 2:
    C = int(sys.argv[2])
p = Problem(arr)
nb = max(3, int(400e6 / (p.D * C * 4)) + 1)
for minor in (True, False):
    shape = (p.D, C) if minor else (C, p.D)
    sets = [(torch.rand(shape, device="cuda") * 4 - 2) for _ in range(nb)]
    grads = [torch.empty_like(s) for s in sets]
    lp = torch.empty(C, device="cuda"); cc = torch.empty(C, device="cuda")
    for i in range(nb):
        p.logdensity(sets[i], chain_minor=minor, lp=lp, grad=grads[i], corr_coef=cc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 30
    e0.record()
    for i in range(reps):
        p.logdensity(sets[i % nb], chain_minor=minor, lp=lp, grad=grads[i % nb], corr_coef=cc)
    e1.record(); e1.synchronize()
    print(json.dumps({"workload": desc, "chains": C, "layout": "chain-minor [D, C]" if minor else "chain-major [C, D]",
                      "ms_per_call": e0.elapsed_time(e1) / reps, "finite": bool(torch.isfinite(lp).all())}), flush=True)
