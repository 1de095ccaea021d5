import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from bpl_next_b200 import DixonColesMatchPredictor, diagnostics as dg
from oracle import datasets
for chains in (64, 1024):
    t0 = time.perf_counter()
    m = DixonColesMatchPredictor().fit(datasets.dummy_data(), num_warmup=500, num_samples=250, mcmc_kwargs={"num_chains": chains})
    torch.cuda.synchronize(); wall = time.perf_counter() - t0
    r = m.nuts_run
    print(chains, "wall %.2f launches %d" % (wall, r.launches), "leapfrog per chain pct", np.percentile(r.num_leapfrog, [0, 50, 90, 99, 100]),
          "step pct", np.percentile(r.step_size, [0, 1, 50, 100]).round(4), "div", r.num_divergent.sum())
    worst = int(np.argmax(r.num_leapfrog))
    print("  worst chain", worst, "step", r.step_size[worst], "accept mean", r.accept[:, worst].mean().item(), "imm min/max", r.inv_mass[:, worst].min().item(), r.inv_mass[:, worst].max().item())
