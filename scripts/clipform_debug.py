"""Which gradient entries differ between the clip-free and the clipping forms (BPLX_CLIP_FORMS forces the latter)."""
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
from tests import helpers as H
from bpl_next_b200 import Problem
arr = H.small_problem("extended", seed=3, weighted=True, K=3)
C = 64
near = H.random_theta(arr_D := 34, C, seed=5, radius=0.15, dtype=np.float32)
res = {}
for forms in ("0", "1", "2", "3"):
    os.environ["BPLX_CLIP_FORMS"] = forms
    p = Problem(arr)
    res[forms] = [x.cpu().numpy() for x in p.logdensity(torch.from_numpy(near).cuda())]
    p.close()
for forms in ("1", "2", "3"):
    d = res[forms][1] != res["0"][1]
    print("forms", forms, "lp equal", np.array_equal(res[forms][0], res["0"][0]), "cc equal", np.array_equal(res[forms][2], res["0"][2]),
          "grad diffs per index", d.sum(axis=0))
