"""Fixed cost of a K1 launch: tiny problems at 4,096 chains (CUDA graph of 200 launches, rotating buffers)."""
import sys, os, numpy as np, torch
sys.path.insert(0, '.')
import bench
from bpl_next_b200 import Problem, data as bdata
from tests import helpers as H
for desc, arr in [("DC T=2 M=2", H.small_problem("dixon_coles", T=2, M=2)),
                  ("DC T=20 M=380", bdata.prepare("dixon_coles", __import__("oracle.datasets", fromlist=["x"]).dummy_data())[0]),
                  ("EXT T=20 M=1900 K=3 (cfg2)", bench.workload("cfg2")[0]),
                  ("EXT T=20 M=1900 K=0", None)]:
    if arr is None:
        from oracle import datasets
        td = datasets.config_2(); td.pop("team_covariates")
        arr = bdata.prepare("extended", td, epsilon=0.01)[0]
    p = Problem(arr)
    ms, reps, nb, sb, fin = bench.time_logdensity(p, 4096, 200, 5, 1.0, 1, use_graph=True, target_s=0.5)
    print(f"{desc:30s} us/launch {1e3*ms/200:.2f}  plan {p.stats()}")
