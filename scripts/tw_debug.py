import sys, numpy as np
sys.path.insert(0, '.')
from bpl_next_b200 import ExtendedDixonColesMatchPredictor
from oracle import datasets
td = datasets.timed_dummy_data()
for seed in (42, 7):
    kw = dict(num_warmup=400, num_samples=200, random_state=seed, mcmc_kwargs={"num_chains": 256})
    res = {}
    for eps in (1, 2):
        m = ExtendedDixonColesMatchPredictor().fit(td, epsilon=eps, **kw)
        a = m.attack.mean(axis=0); d = m.defence.mean(axis=0)
        res[eps] = (a[1] - a[0], d[1] - d[0], m.nuts_run.num_divergent.sum(), m.nuts_run.step_size.mean())
    print(seed, res, "ratio attack", abs(res[2][0]) / abs(res[1][0]))
