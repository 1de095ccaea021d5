"""Per-launch cost of one leapfrog of a configs[2] fit in steady state: CUDA-event times of the blocks of 32 (log-density
kernel + NUTS step kernel) pairs of one run (CUDA graph replays), next to the log-density kernel alone.
usage: python scripts/fit_step_time.py [chains] [launches] [unused] [auto|chain_minor|chain_major]"""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import os
from bpl_next_b200 import _abi
if os.environ.get("BPLX_LIB"):  # a library variant built for an experiment
    _abi.LIB_PATH = os.path.join(os.path.dirname(_abi.LIB_PATH), os.environ["BPLX_LIB"])
from bpl_next_b200 import Problem, nuts as bn, data as bdata
from oracle import datasets

C = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
n1 = int(sys.argv[2]) if len(sys.argv) > 2 else 512
n2 = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
LAYOUT = sys.argv[4] if len(sys.argv) > 4 else "auto"  # auto | chain_minor | chain_major
arr, _ = bdata.prepare("neutral_wc", datasets.config_3(), epsilon=0.1)
p = Problem(arr)
g = torch.Generator(device="cuda").manual_seed(1)
theta0 = (torch.rand((p.D, C), generator=g, device="cuda") * 4 - 2).contiguous()


def potential(theta, lp, grad):
    p.logdensity(theta, chain_minor=True, lp=lp, grad=grad)


def potential_cm(theta, lp, grad):
    p.logdensity(theta, chain_minor=False, lp=lp, grad=grad)


torch.cuda.synchronize()
r = bn.sample(potential, theta0, num_warmup=1000, num_samples=10, max_tree_depth=6, max_launches=n1, check_every=32, diag_lags=8,
              potential_cm=potential_cm, state_layout=LAYOUT, time_blocks=True,
              use_graph=not os.environ.get("BPLX_NO_GRAPH"))  # (plain launches for an ncu capture)
blocks = np.array(r.block_ms) / 32.0
lp = torch.empty(C, device="cuda")
cm = LAYOUT in ("chain_major", "auto")
th = theta0.t().contiguous() if cm else theta0
grad = torch.empty_like(th)
f = potential_cm if cm else potential
for _ in range(3):
    f(th, lp, grad)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    f(th, lp, grad)
e1.record(); e1.synchronize()
k1 = e0.elapsed_time(e1) / 10
half = blocks[len(blocks) // 2:]
pair = float(np.median(half))
print(json.dumps({"chains": C, "D": p.D, "state_layout": "chain_major" if cm else "chain_minor", "launches": r.launches,
                  "ms_per_pair_steady": pair, "pair_min_max": [float(half.min()), float(half.max())], "k1_ms": k1,
                  "step_ms": pair - k1, "ms_per_pair_first_blocks": [float(x) for x in blocks[:3]]}))
