"""One NUTS run (DixonColes dummy season) at the chain count given on the command line -- a target for ncu and for
comparing the two step kernels (BPLX_NUTS_GENERIC=1): seeded, so both do exactly the same launches.
usage: python scripts/fit_one.py [chains] [nograph]"""
import sys, time, torch
sys.path.insert(0, '.')
from bpl_next_b200 import Problem, nuts as N
from oracle import datasets
from tests import helpers as H
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
arr = H.from_training_data("dixon_coles", datasets.dummy_data())
p = Problem(arr)
g = torch.Generator(device="cuda").manual_seed(11)
theta0 = torch.rand((p.D, chains), generator=g, device="cuda") * 4 - 2
def potential(theta, lp, grad):
    p.logdensity(theta, chain_minor=True, lp=lp, grad=grad)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    run = N.sample(potential, theta0.clone(), num_warmup=100, num_samples=50, seed=1, use_graph=len(sys.argv) <= 2)
    torch.cuda.synchronize(); wall = time.perf_counter() - t0
    print(f"chains {chains}: wall {wall:.3f} s, launches {run.launches}, {1e6 * wall / run.launches:.1f} us per (K1 + step)")
