"""Builds the debug library bpl_next_b200/lib/libbplx_timeline.so (K1 with -DBPLX_TIMELINE, see scripts/timeline.py)
and prints the ptxas resource lines of the K1 kernels."""
import os, subprocess, sys
sys.path.insert(0, ".")
from bpl_next_b200 import build as B
csrc = os.path.join(os.path.dirname(os.path.abspath(B.__file__)), "csrc")
out = os.path.join(os.path.dirname(B.OUT), "libbplx_timeline.so")
cmd = [os.environ.get("NVCC", "nvcc")] + B.NVCC_FLAGS + ["-DBPLX_TIMELINE", "-Xptxas", "-v", "-o", out] + B.SOURCES
r = subprocess.run(cmd, capture_output=True, text=True, cwd=csrc)
lines = r.stderr.splitlines()
for i, l in enumerate(lines):
    if "Compiling entry function" in l and "logdensity_kernel" in l:
        print(l[:110]); print(lines[i + 1]); print(lines[i + 2])
if r.returncode:
    print(r.stderr[-4000:])
sys.exit(r.returncode)
