"""include/bplx.h is valid ISO C and the ABI works without ctypes: a plain C11 program (tests/c_abi/abi_drive.c) built with
``gcc -std=c11 -pedantic -Wall -Werror`` drives create -> fwdbwd_host -> score_grid_host -> destroy."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "abi_drive.c")
LIBDIR = os.path.join(ROOT, "bpl_next_b200", "lib")


def _build(tmp_path):
    exe = str(tmp_path / "abi_drive")
    cmd = ["gcc", "-std=c11", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", LIBDIR, "-lbplx", "-lm", f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_headers_are_iso_c_and_program_links(tmp_path):
    for hdr in ("bplx.h", "bplx_nuts.h"):
        r = subprocess.run(["gcc", "-std=c11", "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c",
                            os.path.join(ROOT, "include", hdr)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    # without a GPU the library refuses loudly (no CPU fallback); with one the program runs through
    assert r.returncode in (0, 3), (r.returncode, r.stdout, r.stderr)
    if r.returncode == 3:
        assert "no usable CUDA device" in r.stdout


@pytest.mark.gpu
def test_c_program_numbers_match_the_oracle(tmp_path):
    from bpl_next_b200 import data as bdata
    from oracle import models as om, predict as op
    from tests import helpers as H

    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    out = r.stdout
    T = 5
    pairs = [(h, a) for h in range(T) for a in range(T) if h != a]
    arr = bdata.MatchArrays(model="dixon_coles", num_teams=T,
                            home_team=np.array([h for h, _ in pairs], np.uint16), away_team=np.array([a for _, a in pairs], np.uint16),
                            home_goals=np.array([(h * 3 + a * 5 + 1) % 4 for h, a in pairs], np.uint8),
                            away_goals=np.array([(h + 2 * a) % 3 for h, a in pairs], np.uint8))
    D = 5 + 2 * T
    assert f"D {D} " in out and "attack_decentered" in out and "loglik_inputs 12" in out
    theta = np.array([[0.3 * np.sin(1.0 + d + 0.5 * c) for d in range(D)] for c in range(3)], dtype=np.float32)
    lp, g, cc = om.log_density_and_grad(H.to_oracle(arr), theta.astype(np.float64))
    rows = re.findall(r"chain (\d) lp (\S+) corr_coef (\S+) gradnorm (\S+)", out)
    assert len(rows) == 3
    for c, a, b, n in rows:
        c = int(c)
        assert abs(float(a) - lp[c]) < 1e-5 * abs(lp[c])
        assert abs(float(b) - cc[c]) < 1e-5
        assert abs(float(n) - np.linalg.norm(g[c])) < 1e-4 * np.linalg.norm(g[c])
    s = {"attack": np.array([[0.1 * (t - 2) + 0.05 * k for t in range(T)] for k in range(2)], np.float32),
         "defence": np.array([[0.05 * (2 - t) for t in range(T)] for k in range(2)], np.float32),
         "home_advantage": np.array([0.2, 0.3], np.float32), "corr_coef": np.array([0.05, -0.02], np.float32)}
    grid, _, _ = op.predict_score_grid_proba("dixon_coles", s, np.array([0, 3]), np.array([1, 2]), 5)
    m = re.search(r"grid00 (\S+) (\S+) outcome0 (\S+) (\S+) (\S+)", out)
    assert abs(float(m.group(1)) - grid[0, 0, 0]) < 1e-6 and abs(float(m.group(2)) - grid[1, 0, 0]) < 1e-6
    assert "bad_fixture rc -1" in out
