"""Where the host-buffer call of configs[1] spends its time: the copies and the kernel timed separately (CUDA events),
next to the whole call (wall clock)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from bpl_next_b200 import Problem
arr, C, desc = bench.workload("cfg2")
p = Problem(arr)
D = p.D
th_h = torch.empty((C, D), dtype=torch.float32).pin_memory(); th_h.uniform_(-1, 1)
gr_h = torch.empty((C, D), dtype=torch.float32).pin_memory()
lp_h = torch.empty(C, dtype=torch.float32).pin_memory(); cc_h = torch.empty(C, dtype=torch.float32).pin_memory()
th_d = torch.empty((C, D), device="cuda"); gr_d = torch.empty((C, D), device="cuda")
def ev(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n
print("H2D %d KB: %.1f us" % (C * D * 4 // 1024, ev(lambda: th_d.copy_(th_h, non_blocking=True))))
print("D2H %d KB: %.1f us" % (C * D * 4 // 1024, ev(lambda: gr_h.copy_(gr_d, non_blocking=True))))
print("D2H 16 KB: %.1f us" % ev(lambda: lp_h.copy_(gr_d[0, :1].expand(C).contiguous()[:C] if False else gr_d.view(-1)[:C], non_blocking=True)))
print("K1 (device resident, eager back to back): %.1f us" % ev(lambda: p.logdensity(th_d)))
def call():
    p.logdensity_host(th_h.numpy(), lp=lp_h.numpy(), grad=gr_h.numpy(), corr_coef=cc_h.numpy())
for _ in range(10): call()
t0 = time.perf_counter()
for _ in range(300): call()
print("host call: %.1f us wall" % (1e6 * (time.perf_counter() - t0) / 300))
