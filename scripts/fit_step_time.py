"""Per-launch cost of one leapfrog of a configs[2] fit: the log-density kernel (K1) and the NUTS step kernel, timed with
CUDA events over a fixed number of (K1, step) pairs (no CUDA graph, so that ncu can pick single launches)."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from bpl_next_b200 import Problem, nuts as bn, data as bdata
from oracle import datasets

C = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 96
arr, _ = bdata.prepare("neutral_wc", datasets.config_3(), epsilon=0.1)
p = Problem(arr)
g = torch.Generator(device="cuda").manual_seed(1)
theta0 = (torch.rand((p.D, C), generator=g, device="cuda") * 4 - 2).contiguous()
t_k1 = []


def potential(theta, lp, grad):
    p.logdensity(theta, chain_minor=True, lp=lp, grad=grad)


torch.cuda.synchronize()
t0 = time.perf_counter()
run = bn.sample(potential, theta0, num_warmup=1000, num_samples=10, max_tree_depth=6, max_launches=pairs, use_graph=False,
                check_every=32, diag_lags=8)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
# K1 alone on the same shapes
lp = torch.empty(C, device="cuda"); grad = torch.empty_like(theta0)
for _ in range(3):
    potential(theta0, lp, grad)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    potential(theta0, lp, grad)
e1.record(); e1.synchronize()
k1 = e0.elapsed_time(e1) / 10
print(json.dumps({"chains": C, "D": p.D, "pairs": run.launches, "wall_s": wall, "ms_per_pair": 1e3 * wall / run.launches,
                  "k1_ms": k1, "step_ms_est": 1e3 * wall / run.launches - k1}))
