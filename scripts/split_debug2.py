import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from bpl_next_b200 import Problem
from oracle import datasets
from tests import helpers as H
arr = H.from_training_data("neutral_wc", datasets.config_3(), epsilon=0.1)
base = H.random_theta(1339, 32, seed=21, radius=0.3, dtype=np.float32)
res = {}
for raw in (-30.0, 30.0):
    theta = base.copy(); theta[:, -1] = raw   # r -> 0: cc = LB ; r -> 1: cc = UB
    for split in (1, 2, 8):
        os.environ["BPLX_SPLIT"] = str(split)
        p = Problem(arr)
        lp, g, cc = p.logdensity(torch.from_numpy(theta).cuda()); torch.cuda.synchronize()
        res[(raw, split)] = cc.cpu().numpy().copy()
        p.close()
    for split in (2, 8):
        d = res[(raw, split)] - res[(raw, 1)]
        print("raw", raw, "split", split, "chains differing", int((np.abs(d) > 1e-6).sum()), "of 32; first vals", res[(raw, 1)][:4], res[(raw, split)][:4])
