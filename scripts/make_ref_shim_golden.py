"""Golden vectors from the reference's own SOURCE (run in the build container, where /root/reference exists).

For each case: the reference class's real `fit()` is called on a data set of oracle/datasets.py (the reference's conftest
fixtures regenerated with the same seeds), `MCMC.run` is intercepted (oracle/ref_shim.py), and the captured `_model` is
traced at random unconstrained positions: log joint density and gradient per site.  Stored in this repo's flat layout
(oracle/models.py `site_layout`: numpyro's site names) as tests/golden/ref_shim_<case>.npz together with the positions.
tests/test_golden.py::test_oracle_matches_reference_source compares the oracle restatement with them."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H  # (before the reference is put on sys.path: it has a `tests` package too)
from oracle import ref_shim
ref_shim.install()
import bpl  # the reference, unmodified, under the stand-ins
from bpl.dynamic_dixon_coles import DynamicNeutralDixonColesMatchPredictor as RefDynamic  # noqa: F401 (import check)
from oracle import datasets, models as om
from bpl_next_b200 import data as bdata

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


REF_CLASS = {"dixon_coles": bpl.DixonColesMatchPredictor, "extended": bpl.ExtendedDixonColesMatchPredictor,
             "neutral": bpl.NeutralDixonColesMatchPredictor, "neutral_wc": bpl.NeutralDixonColesMatchPredictorWC}
# (the dynamic class is left out: as written its walk never reaches the rates and it indexes a G-row array with gameweek
#  G, which jax clamps silently -- SURVEY D1, D2; the oracle's own tests cover both readings)
CASES = {name: (model, REF_CLASS[model], td, kw) for name, (model, td, kw) in datasets.ref_shim_cases().items()}

for case, (model, cls, td, kw) in CASES.items():
    cap = ref_shim.capture_fit(cls(), td, **kw)
    arr, meta = bdata.prepare(model, td, epsilon=kw.get("epsilon"), rescale_weights=kw.get("rescale_weights", False))
    K = 0 if arr.covariates is None else arr.covariates.shape[1]
    layout = om.site_layout(model, arr.num_teams, K, arr.num_conferences or 0, 0)
    offs = {k: v for k, v in om.layout_offsets(layout).items() if k != "__D__"}
    D = om.num_params(model, arr.num_teams, K, arr.num_conferences or 0, 0)
    rng = np.random.default_rng(7)
    radii = np.array([0.3, 0.6, 1.0, 1.0, 1.5, 2.0]) if D < 200 else np.array([0.3, 1.0, 2.0])  # (the big configs: three)
    n = len(radii)
    theta = rng.uniform(-1.0, 1.0, (n, D)) * radii[:, None]
    lps, grads, dets = [], [], {}
    for i in range(n):
        vals = {}
        for name, (o, shape, _tr) in offs.items():
            cnt = int(np.prod(shape)) if shape else 1
            if cnt == 0:
                continue
            vals[name] = theta[i, o:o + cnt].reshape(shape)
        lp, g, det = ref_shim.log_density(cap, vals)
        flat = np.zeros(D)
        for name, (o, shape, _tr) in offs.items():
            cnt = int(np.prod(shape)) if shape else 1
            if cnt:
                flat[o:o + cnt] = np.asarray(g[name]).reshape(-1)
        lps.append(lp)
        grads.append(flat)
        for k, v in det.items():  # the deterministic sites the reference's fit() reads back from the samples
            dets.setdefault(k, []).append(np.asarray(v))
    np.savez_compressed(os.path.join(OUT, f"ref_shim_{case}.npz"), theta=theta, lp=np.array(lps), grad=np.array(grads),
                        model=model, kwargs=repr(kw), **{"det_" + k: np.array(v) for k, v in dets.items()})
    # the oracle on the same positions, right away (the committed test repeats this without the reference)
    lo, go, _ = om.log_density_and_grad(H.to_oracle(arr), theta)
    err_lp = np.max(np.abs(lo - np.array(lps)) / np.maximum(1.0, np.abs(np.array(lps))))
    err_g = np.max(np.abs(go - np.array(grads)) / np.maximum(1.0, np.abs(np.array(grads)).max(axis=1, keepdims=True)))
    print(f"{case:24s} D={D:4d}  lp ref {lps[0]:.6f} oracle {lo[0]:.6f}   max rel err lp {err_lp:.2e} grad {err_g:.2e}")
