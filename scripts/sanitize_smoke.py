"""Small invocations of every kernel (K1 all models, K1d, K3, NUTS) for compute-sanitizer."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from bpl_next_b200 import Problem, score_grid, nuts as bn
from oracle import datasets
from tests import helpers as H
for model, kw in [("dixon_coles", {}), ("extended", dict(K=3)), ("neutral", {}), ("neutral_wc", dict(multi_conf=True, K=2)),
                  ("dynamic", dict(Cf=4, T=5, M=90))]:
    arr = H.small_problem(model, seed=3, **kw)
    p = Problem(arr)
    for C, minor in ((45, False), (33, True)):
        th = torch.from_numpy(H.random_theta(p.D, C, seed=1, dtype=np.float32)).cuda()
        if minor: th = th.t().contiguous()
        lp, g, cc = p.logdensity(th, chain_minor=minor)
        torch.cuda.synchronize()
        assert torch.isfinite(lp).all()
    p.close()
s, fx = datasets.config_5(S=64, F=300)
ds = {k: torch.from_numpy(v).cuda() for k, v in s.items()}; dfx = {k: torch.from_numpy(v).cuda() for k, v in fx.items()}
for mg in (10, 15):
    grid, out = score_grid("neutral_wc", ds, dfx, mg); torch.cuda.synchronize()
arr = H.from_training_data("dixon_coles", datasets.dummy_data()); p = Problem(arr)
def potential(theta, lp, grad): p.logdensity(theta, chain_minor=True, lp=lp, grad=grad)
run = bn.sample(potential, torch.rand((p.D, 40), device="cuda") * 2 - 1, num_warmup=30, num_samples=10, use_graph=False, check_every=8)
torch.cuda.synchronize(); print("sanitize smoke ok", run.launches)
