"""CPU: the C-ABI library loads and exports every symbol include/bplx.h declares; no compute calls."""
import ctypes as C
import os
import re

from bpl_next_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bplx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bplx_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = _abi.lib()
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/bplx.h but not exported by libbplx.so"


def test_version_and_error_string():
    lib = _abi.lib()
    assert lib.bplx_version() == 1
    assert isinstance(lib.bplx_last_error(), bytes)
    assert lib.bplx_launch_count() >= 0


def test_struct_layouts_match_header():
    # field order / sizes of the ctypes mirrors of bplx_problem_desc, bplx_samples, bplx_fixtures
    assert C.sizeof(_abi.ProblemDesc) == 7 * 4 + 4 + 10 * 8  # 7 int32/uint32, padding, 10 pointers
    assert C.sizeof(_abi.Samples) == 4 * 4 + 8 * 8
    assert C.sizeof(_abi.Fixtures) == 8 + 5 * 8


def test_invalid_arguments_fail_cleanly_without_a_gpu():
    lib = _abi.lib()
    h = C.c_void_p()
    assert lib.bplx_problem_create(None, C.byref(h)) == _abi.E_INVALID
    assert b"NULL" in lib.bplx_last_error()
    assert lib.bplx_num_params(None) == _abi.E_INVALID
    assert lib.bplx_logdensity_workspace_bytes(None, 8) == 0
    lib.bplx_problem_destroy(None)  # no-op
