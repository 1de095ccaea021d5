"""Repeat K1 on the same input and count differing output bits (models with and without rate clipping)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from bpl_next_b200 import Problem
from oracle import datasets
from tests import helpers as H
for name, arr in (("dixon_coles dummy", H.from_training_data("dixon_coles", datasets.dummy_data())),
                  ("extended small", H.small_problem("extended", seed=3, weighted=True, K=3)),
                  ("neutral small", H.small_problem("neutral", seed=3, K=2))):
    p = Problem(arr)
    for radius in (0.3, 2.0):
        t = torch.from_numpy(H.random_theta(p.D, 1024, seed=3, radius=radius, dtype=np.float32)).cuda().t().contiguous()
        ref = [x.clone() for x in p.logdensity(t, chain_minor=True)]
        bad = [0, 0, 0]
        for rep in range(30):
            out = p.logdensity(t, chain_minor=True)
            torch.cuda.synchronize()
            for i in range(3):
                bad[i] += int((out[i] != ref[i]).sum().item())
        print(name, "radius", radius, "differing lp / grad / cc entries over 30 repeats:", bad)
    p.close()
