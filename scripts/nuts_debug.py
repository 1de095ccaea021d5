import sys, numpy as np, torch
sys.path.insert(0, '.')
from bpl_next_b200 import nuts as bn, diagnostics as dg
D, C = 6, 512
mu = torch.tensor([0.0, 1.0, -2.0, 3.0, 0.5, -0.5], device="cuda")[:, None]
sd = torch.tensor([1.0, 0.1, 10.0, 2.0, 0.5, 5.0], device="cuda")[:, None]
def potential(theta, lp, grad):
    z = (theta - mu) / sd
    lp.copy_(-0.5 * (z * z).sum(0))
    grad.copy_(-z / sd)
g = torch.Generator(device="cuda").manual_seed(0)
theta0 = (torch.rand((D, C), generator=g, device="cuda") * 4 - 2)
for nw in (300, 1000):
    run = bn.sample(potential, theta0, num_warmup=nw, num_samples=200, seed=1)
    x = run.samples.double()
    print("warmup", nw, "launches", run.launches)
    print(" mean", x.mean(dim=(0, 2)).cpu().numpy().round(3))
    print(" std ", x.permute(1, 0, 2).reshape(D, -1).std(dim=1).cpu().numpy().round(3))
    print(" imm mean", run.inv_mass.mean(1).cpu().numpy().round(3), "imm min", run.inv_mass.min(1).values.cpu().numpy().round(4))
    print(" step", np.percentile(run.step_size, [0, 50, 100]).round(3), "accept", run.accept.mean().item(), "leapfrog/transition", run.num_leapfrog.mean() / (nw + 200))
    print(" ess", dg.effective_sample_size(run.samples).cpu().numpy().round(0), "rhat", dg.split_rhat(run.samples).cpu().numpy().round(3))
    # per-chain mean of dim 2
    m2 = x[:, 2, :].mean(0)
    print(" chain means dim2: mean %.3f sd %.3f (expected sd %.3f)" % (m2.mean().item(), m2.std().item(), 10 / np.sqrt(200)))
