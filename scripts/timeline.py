"""Per-warp region timeline of K1 on CTA 0 (debug library built with -DBPLX_TIMELINE, see logdensity.cu).
usage: python scripts/timeline.py [cfg2|cfg3|cfg1] [radius] [chains]"""
import ctypes, os, sys, numpy as np, torch
sys.path.insert(0, '.')
from bpl_next_b200 import _abi
_abi.LIB_PATH = os.path.join(os.path.dirname(_abi.LIB_PATH), "libbplx_timeline.so")
import bench
from bpl_next_b200 import Problem
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
radius = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0  # theta ~ U(-radius, radius)
arr, C, desc = bench.workload(wl)
if len(sys.argv) > 3:
    C = int(sys.argv[3])
p = Problem(arr)
lib = _abi.lib()
sets = [(torch.rand((p.D, C), device="cuda") * 2 - 1) * radius for _ in range(60)]  # rotate: theta misses L2 like in the bench
names = ["start", "prologue done", "after barrier P", "phase 1 done", "after bounds barrier", "search done", "phase 2 done",
         "after barrier B+gc", "fix-up done", "team pass done", "after partial barrier", "epilogue done",
         "(prologue) first theta load back", "(prologue) hyper + first team loads back", "(bounds) after max barrier", "(bounds) after offset barrier", "(phase 1) first stage acquired", "(phase 1) first piece done", "(phase 2) first stage acquired", "kernel entry"]
acc = None
for it in range(len(sets)):
    p.logdensity(sets[it], chain_minor=True)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (32 * 24))()
    lib.bplx_timeline_read.argtypes = [ctypes.c_void_p]
    assert lib.bplx_timeline_read(buf) == 0
    t = np.array(buf, dtype=np.uint64).reshape(32, 24)[: p.stats()["warps"], :20].astype(np.float64)
    t -= t[:, 19].min()
    if it >= 10:
        acc = t if acc is None else acc + t
acc /= (len(sets) - 10)
print(desc, "warps", acc.shape[0], "radius", radius)
for i, n in enumerate(names):
    if abs(acc[:, i]).max() > 1e12:  # (a stamp this build does not take)
        continue
    print(f"{n:24s} min {acc[:, i].min() / 1e3:7.2f} us   max {acc[:, i].max() / 1e3:7.2f} us")

# per-warp phase times against what the plan dealt to each warp: least-squares cost per piece / entry (plan.cc's model)
W = acc.shape[0]
buf = (ctypes.c_longlong * (W * 12))()
if lib.bplx_problem_warp_stats(p._h, buf, W * 12) == W * 12:
    st = np.array(buf, dtype=np.float64).reshape(W, 12)
    t1 = (acc[:, 3] - acc[:, 2]) * 1.98  # cycles at 1.98 GHz (ns -> cycles)
    t2 = (acc[:, 6] - acc[:, 5]) * 1.98
    for name, t, A, cols in (("phase 1", t1, st[:, :6], ["home pieces", "away pieces", "home entries", "away entries", "teams", "stages"]),
                             ("phase 2", t2, st[:, 6:], ["pieces", "XY entries", "X entries", "Y entries", "teams", "stages"])):
        keep = [i for i in range(6) if A[:, i].std() > 0]
        X = np.concatenate([A[:, keep], np.ones((W, 1))], axis=1)
        coef, *_ = np.linalg.lstsq(X, t, rcond=None)
        print(name, "cycles per", {cols[k]: round(float(c), 1) for k, c in zip(keep, coef[:-1])}, "const", round(float(coef[-1])),
              "residual rms", round(float(np.sqrt(np.mean((X @ coef - t) ** 2)))), "of mean", round(float(t.mean())))
        for w in range(W):
            print(f"  warp {w:2d}  {t[w]:9.0f} cycles ", " ".join(f"{int(x):6d}" for x in A[w]))
