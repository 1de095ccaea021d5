"""K1 on one far-out position (tests/golden/far_out_theta_dixon_coles.npy, found by scripts/nuts_determinism.py): is the output the same
on every call?"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from bpl_next_b200 import Problem
from oracle import datasets, models as om
from tests import helpers as H
arr = H.from_training_data("dixon_coles", datasets.dummy_data())
p = Problem(arr)
th = np.load("tests/golden/far_out_theta_dixon_coles.npy").astype(np.float32)
print("layout", p.layout)
good = H.random_theta(p.D, 64, seed=1, radius=1.0, dtype=np.float32)
good[5] = th
t = torch.from_numpy(good).cuda()
outs = []
for rep in range(40):
    lp, grad, cc = p.logdensity(t)
    torch.cuda.synchronize()
    outs.append((lp[5].item(), cc[5].item(), grad[5].cpu().numpy().copy()))
print("lp values", sorted(set(o[0] for o in outs)), "cc", sorted(set(str(o[1]) for o in outs)))
g0 = outs[0][2]
nd = sum(1 for o in outs if not np.array_equal(np.nan_to_num(o[2], nan=123.0), np.nan_to_num(g0, nan=123.0)))
print("calls whose gradient differs from the first:", nd)
print("grad first", np.round(g0, 2))
for o in outs[1:]:
    if not np.array_equal(np.nan_to_num(o[2], nan=123.0), np.nan_to_num(g0, nan=123.0)):
        print("grad other", np.round(o[2], 2)); break
lo, go, co = om.log_density_and_grad(H.to_oracle(arr), th[None].astype(np.float64))
print("oracle float64 lp", lo, "cc", co, "grad finite", np.isfinite(go).all())
lo, go, co = om.log_density_and_grad(H.to_oracle(arr), th[None], dtype=torch.float32)
print("oracle float32 lp", lo, "cc", co, "grad finite", np.isfinite(go).all())
