"""Where a NUTS run spends its time: wall per launch pair (K1 + step) at a few chain counts, then the two kernels alone
(CUDA events around repeated eager launches)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from bpl_next_b200 import Problem, nuts as N
from oracle import datasets
from tests import helpers as H
arr = H.from_training_data("dixon_coles", datasets.dummy_data())
p = Problem(arr)
for chains in (1, 32, 1024, 4096):
    theta0 = torch.rand((p.D, chains), device="cuda") * 4 - 2
    def potential(theta, lp, grad):
        p.logdensity(theta, chain_minor=True, lp=lp, grad=grad)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    run = N.sample(potential, theta0, num_warmup=200, num_samples=100, seed=1)
    torch.cuda.synchronize(); wall = time.perf_counter() - t0
    print(f"chains {chains}: wall {wall:.3f} s, launches {run.launches}, {1e6 * wall / run.launches:.1f} us per (K1 + step), "
          f"leapfrogs {int(run.num_leapfrog.sum())}")
    # K1 alone
    th = theta0.clone(); lp = torch.zeros(chains, device="cuda"); gr = torch.zeros_like(th)
    for _ in range(5): potential(th, lp, gr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(50): potential(th, lp, gr)
        g.replay(); s.synchronize()
        e0.record(s); g.replay(); e1.record(s); e1.synchronize()
    print(f"   K1 alone: {1e3 * e0.elapsed_time(e1) / 50:.1f} us per launch")
