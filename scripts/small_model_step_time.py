"""Leapfrog cost of small-model fits per sampler state layout: chain-minor (register-resident step kernels, lane = chain)
against chain-major (a warp per chain).  CUDA events per block of 32 (log-density, step) pairs of a real run.
usage: python scripts/small_model_step_time.py"""
import json, sys
import numpy as np, torch
sys.path.insert(0, '.')
from bpl_next_b200 import Problem, nuts as bn, data as bdata
from oracle import datasets

for name, model, td, kw, C in (("configs[0] DixonColes T=20 M=380", "dixon_coles", datasets.dummy_data(), {}, 1024),
                               ("configs[0] DixonColes T=20 M=380", "dixon_coles", datasets.dummy_data(), {}, 1),
                               ("configs[1] Extended T=20 M=1900 K=3", "extended", datasets.config_2(), dict(epsilon=0.01), 4096)):
    arr, _ = bdata.prepare(model, td, **kw)
    p = Problem(arr)
    g = torch.Generator(device="cuda").manual_seed(1)
    theta0 = (torch.rand((p.D, C), generator=g, device="cuda") * 4 - 2).contiguous()

    def potential(theta, lp, grad):
        p.logdensity(theta, chain_minor=True, lp=lp, grad=grad)

    def potential_cm(theta, lp, grad):
        p.logdensity(theta, chain_minor=False, lp=lp, grad=grad)

    for lay in ("chain_minor", "chain_major"):
        r = bn.sample(potential, theta0, num_warmup=300, num_samples=50, max_launches=4096, check_every=32,
                      potential_cm=potential_cm, state_layout=lay, time_blocks=True)
        b = np.array(r.block_ms) / 32.0
        half = b[len(b) // 2:]
        print(json.dumps({"workload": name, "chains": C, "D": p.D, "state_layout": lay, "us_per_leapfrog": 1e3 * float(np.median(half)),
                          "min_max": [1e3 * float(half.min()), 1e3 * float(half.max())]}), flush=True)
    p.close()
