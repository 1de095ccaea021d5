"""Few-chain regime: per-call time of the log-density entry point at small chain counts (K1's cluster plans), on configs[2],
configs[1] and configs[0] data (CUDA graph of 50 calls, CUDA events).  With an argument: only that many chains of configs[2]
data, 20 plain launches (for an ncu capture of the single-chain call)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from bpl_next_b200 import Problem, _abi, data as bdata
from oracle import datasets

out = []
for name, model, td, kw in (("configs[2] data T=220 M=40000", "neutral_wc", datasets.config_3(), dict(epsilon=0.1)),
                            ("configs[1] data T=20 M=1900", "extended", datasets.config_2(), dict(epsilon=0.01)),
                            ("configs[0] data T=20 M=380", "dixon_coles", datasets.dummy_data(), {})):
    arr, _ = bdata.prepare(model, td, **kw)
    for few in (0,):
        p = Problem(arr)
        if len(sys.argv) > 1:
            C = int(sys.argv[1])
            th = torch.rand((p.D, C), device="cuda") * 2 - 1
            for _ in range(20):
                out_ = p.logdensity(th, chain_minor=True)
            torch.cuda.synchronize()
            print(json.dumps({"data": name, "chains": C, "finite": bool(torch.isfinite(out_[0]).all())}))
            sys.exit(0)
        for C in (1, 8, 32, 148, 296, 592, 1024):
            g = torch.Generator(device="cuda").manual_seed(C)
            th = (torch.rand((p.D, C), generator=g, device="cuda") * 2 - 1)
            lp = torch.empty(C, device="cuda"); gr = torch.empty_like(th); cc = torch.empty(C, device="cuda")
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(3):
                    p.logdensity(th, chain_minor=True, lp=lp, grad=gr, corr_coef=cc, stream=s)
                s.synchronize()
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=s):
                    for _ in range(50):
                        p.logdensity(th, chain_minor=True, lp=lp, grad=gr, corr_coef=cc, stream=s)
                gph.replay(); s.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ts = []
                for _ in range(5):
                    e0.record(s); gph.replay(); e1.record(s); e1.synchronize(); ts.append(e0.elapsed_time(e1) / 50)
            out.append({"data": name, "kernel": "K1", "chains": C, "us_per_call": 1e3 * float(np.median(ts)),
                        "finite": bool(torch.isfinite(lp).all())})
            print(json.dumps(out[-1]), flush=True)
        p.close()
