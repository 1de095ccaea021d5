import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from bpl_next_b200 import Problem
from oracle import datasets, models as om
from tests import helpers as H
arr = H.from_training_data("neutral_wc", datasets.config_3(), epsilon=0.1)
theta = H.random_theta(1339, 40, seed=21, radius=0.3, dtype=np.float32)
lp_o, g_o, cc_o = om.log_density_and_grad(H.to_oracle(arr), theta.astype(np.float64))
for split in (1, 2, 4, 8):
    os.environ["BPLX_SPLIT"] = str(split)
    p = Problem(arr)
    lp, g, cc = p.logdensity(torch.from_numpy(theta).cuda()); torch.cuda.synchronize()
    lp, g, cc = lp.cpu().numpy(), g.cpu().numpy(), cc.cpu().numpy()
    sc = np.abs(g_o).max(1, keepdims=True)
    err = np.abs(g - g_o) / sc
    print(split, "lp rel err %.2e" % np.abs(lp / lp_o - 1).max(), "cc abs err %.2e" % np.abs(cc - cc_o).max(), "grad scaled err %.2e at" % err.max(), np.unravel_index(err.argmax(), err.shape), p.stats()["smem_bytes"])
    p.close()
