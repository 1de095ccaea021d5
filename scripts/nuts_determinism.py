"""Is K1 reproducible on the positions a NUTS run visits?  Every potential call is made twice and compared."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from bpl_next_b200 import Problem, nuts as N
from oracle import datasets
from tests import helpers as H
arr = H.from_training_data("dixon_coles", datasets.dummy_data())
p = Problem(arr)
chains = 256
g = torch.Generator(device="cuda").manual_seed(11)
theta0 = torch.rand((p.D, chains), generator=g, device="cuda") * 4 - 2
lp2 = torch.zeros(chains, device="cuda"); gr2 = torch.zeros((p.D, chains), device="cuda")
bad = {"calls": 0, "lp": 0, "grad": 0, "first": None}
def potential(theta, lp, grad):
    p.logdensity(theta, chain_minor=True, lp=lp, grad=grad)
    p.logdensity(theta, chain_minor=True, lp=lp2, grad=gr2)
    dl = ~((lp == lp2) | (torch.isnan(lp) & torch.isnan(lp2)))
    dg = ~((grad == gr2) | (torch.isnan(grad) & torch.isnan(gr2)))
    bad["calls"] += 1
    nl, ng = int(dl.sum().item()), int(dg.sum().item())
    bad["lp"] += nl; bad["grad"] += ng
    if (nl or ng) and bad["first"] is None:
        c = int(torch.nonzero(dg.any(dim=0) | dl)[0].item())
        bad["first"] = (bad["calls"], c, theta[:, c].clone(), lp[c].item(), lp2[c].item(), grad[:, c].clone(), gr2[:, c].clone())
run = N.sample(potential, theta0.clone(), num_warmup=60, num_samples=20, seed=1, use_graph=False)
print("calls", bad["calls"], "differing lp", bad["lp"], "grad entries", bad["grad"])
if bad["first"]:
    k, c, th, a, b, ga, gb = bad["first"]
    print("first at call", k, "chain", c, "lp", a, b)
    print("theta", th.cpu().numpy().round(3))
    d = torch.nonzero(ga != gb).flatten().cpu().numpy()
    print("differing grad idx", d, ga[d].cpu().numpy(), gb[d].cpu().numpy())
    np.save("gpurun_out/bad_theta.npy", th.cpu().numpy())
