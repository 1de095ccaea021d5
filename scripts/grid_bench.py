"""Times bplx_score_grid on BASELINE configs[4] (S=16384, F=10000, 11x11) -- used for ncu captures of K3."""
import sys, json, numpy as np, torch
sys.path.insert(0, '.')
from bpl_next_b200 import score_grid
from oracle import datasets
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
s, fx = datasets.config_5(S=S)
ds = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
dfx = {k: torch.from_numpy(v).cuda() for k, v in fx.items()}
F = len(fx["home_team"])
ws = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
grid = torch.empty((F, 11, 11), dtype=torch.float32, device="cuda"); outc = torch.empty((F, 3), dtype=torch.float32, device="cuda")
for _ in range(3): score_grid("neutral_wc", ds, dfx, 10, workspace=ws, grid=grid, outcome=outc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(5):
    e0.record(); score_grid("neutral_wc", ds, dfx, 10, workspace=ws, grid=grid, outcome=outc); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
print(json.dumps({"S": S, "F": F, "ms": float(np.median(ts)), "sum_err": float((outc.sum(1) - 1).abs().max())}))
