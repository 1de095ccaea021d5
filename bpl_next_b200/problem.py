"""Python handles over the C ABI: ``Problem`` (log-density + gradient) and ``score_grid``.

torch is used only for device memory and streams; every number comes from ``libbplx.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _abi
from .data import MatchArrays

_STAT_NAMES = ("matches", "entries1", "entries1_padded", "entries2", "entries2_padded", "smem_bytes", "warps", "vteams")


def _stream_ptr(stream=None) -> int:
    s = stream if stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)


def _dptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else int(t.data_ptr())


class Problem:
    """Static match data bound to the current CUDA device (``bplx_problem``)."""

    def __init__(self, arrays: MatchArrays):
        if not torch.cuda.is_available():
            raise RuntimeError("bpl_next_b200 needs a CUDA device: there is no CPU fallback")
        self._lib = _abi.lib()
        self.arrays = arrays
        self._h = C.c_void_p()
        desc = arrays.desc()
        torch.cuda.current_device()  # make sure the primary context exists
        _abi.check(self._lib.bplx_problem_create(C.byref(desc), C.byref(self._h)))
        self.D = int(self._lib.bplx_num_params(self._h))
        self.layout = self._parse_layout(self._lib.bplx_problem_layout(self._h).decode())
        self._ws: Optional[torch.Tensor] = None

    @staticmethod
    def _parse_layout(s: str) -> Dict[str, Tuple[int, int, str]]:
        out = {}
        for rec in filter(None, s.split(";")):
            name, off, cnt, tr = rec.split(":")
            out[name] = (int(off), int(cnt), tr)
        return out

    def stats(self) -> Dict[str, int]:
        buf = (C.c_longlong * 8)()
        self._lib.bplx_problem_stats(self._h, buf, 8)
        return dict(zip(_STAT_NAMES, [int(x) for x in buf]))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.bplx_problem_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- device path -------------------------------------------------------------------------
    def workspace(self, num_chains: int) -> Optional[torch.Tensor]:
        need = int(self._lib.bplx_logdensity_workspace_bytes(self._h, num_chains))
        if need == 0:
            return None
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device="cuda")
        return self._ws

    def logdensity(self, theta: torch.Tensor, chain_minor: bool = False, lp=None, grad=None, corr_coef=None,
                   stream=None):
        """Enqueue log-density + gradient.  ``theta`` is ``[C, D]`` (chain-major) or ``[D, C]``
        (``chain_minor=True``), float32, on the GPU; rows may be padded (a view with unit inner stride: the row pitch is
        passed as the leading dimension, and ``grad`` must have the same pitch).  Returns ``(lp, grad, corr_coef)``."""
        assert theta.is_cuda and theta.dtype == torch.float32 and theta.dim() == 2
        assert theta.stride(1) == 1 or theta.shape[1] == 1  # (the stride of a one-element dimension means nothing)
        if chain_minor:
            D, Cn = theta.shape
        else:
            Cn, D = theta.shape
        if D != self.D:
            raise ValueError(f"theta has {D} parameters, the model has {self.D}")
        lp = torch.empty(Cn, dtype=torch.float32, device=theta.device) if lp is None else lp
        grad = torch.empty_strided(theta.shape, theta.stride(), dtype=torch.float32, device=theta.device) if grad is None else grad
        assert grad.shape == theta.shape and grad.stride() == theta.stride()
        corr_coef = torch.empty(Cn, dtype=torch.float32, device=theta.device) if corr_coef is None else corr_coef
        ws = self.workspace(Cn)
        ld = int(theta.stride(0)) if theta.shape[0] > 1 else 0
        _abi.check(self._lib.bplx_logdensity_fwdbwd(
            self._h, Cn, _abi.CHAIN_MINOR if chain_minor else _abi.CHAIN_MAJOR, ld,
            _dptr(theta), _dptr(lp), _dptr(grad), _dptr(corr_coef),
            _dptr(ws), 0 if ws is None else ws.numel(), _stream_ptr(stream)))
        return lp, grad, corr_coef

    # ---- likelihood-only entry (what a numpyro model that keeps its prior sites hands to numpyro.factor) -------
    @property
    def loglik_layout(self) -> Dict[str, Tuple[int, int, str]]:
        return self._parse_layout(self._lib.bplx_loglik_layout(self._h).decode())

    def loglik(self, tables: torch.Tensor, chain_minor: bool = False, stream=None):
        """``tables``: the constrained per-team tables packed per ``loglik_layout`` (``[C, Dl]`` or ``[Dl, C]``).
        Returns ``(loglik [C], grad (like tables), corr_coef [C])``: Poisson + tau terms only, no priors."""
        assert tables.is_cuda and tables.dtype == torch.float32 and tables.is_contiguous()
        Dl = int(self._lib.bplx_loglik_num_inputs(self._h))
        if Dl < 0:
            _abi.check(Dl)
        Cn = tables.shape[1] if chain_minor else tables.shape[0]
        if (tables.shape[0] if chain_minor else tables.shape[1]) != Dl:
            raise ValueError(f"tables have {tables.shape} entries, the likelihood takes {Dl} per chain")
        ll = torch.empty(Cn, dtype=torch.float32, device=tables.device)
        grad = torch.empty_like(tables)
        cc = torch.empty(Cn, dtype=torch.float32, device=tables.device)
        ws = self.workspace(Cn)
        _abi.check(self._lib.bplx_loglik_fwdbwd(
            self._h, Cn, _abi.CHAIN_MINOR if chain_minor else _abi.CHAIN_MAJOR, 0, _dptr(tables), _dptr(ll),
            _dptr(grad), _dptr(cc), _dptr(ws), 0 if ws is None else ws.numel(), _stream_ptr(stream)))
        return ll, grad, cc

    # ---- host path (the reference-facing call: numpy in, numpy out) --------------------------------
    def logdensity_host(self, theta: np.ndarray, lp=None, grad=None, corr_coef=None):
        theta = np.ascontiguousarray(theta, dtype=np.float32)
        Cn, D = theta.shape
        if D != self.D:
            raise ValueError(f"theta has {D} parameters, the model has {self.D}")
        lp = np.empty(Cn, dtype=np.float32) if lp is None else lp
        grad = np.empty((Cn, D), dtype=np.float32) if grad is None else grad
        corr_coef = np.empty(Cn, dtype=np.float32) if corr_coef is None else corr_coef
        _abi.check(self._lib.bplx_logdensity_fwdbwd_host(self._h, Cn, theta.ctypes.data, lp.ctypes.data,
                                                         grad.ctypes.data, corr_coef.ctypes.data))
        return lp, grad, corr_coef


# ------------------------------------------------------------------------------------------------------
_SAMPLE_KEYS = ("attack", "defence", "home_attack", "away_attack", "home_defence", "away_defence",
                "confederation_strength", "corr_coef")
_FIXTURE_KEYS = (("home_team", np.uint16), ("away_team", np.uint16), ("home_conf", np.uint8),
                 ("away_conf", np.uint8), ("neutral_venue", np.uint8))


def _samples_struct(model: str, samples: dict, ptr):
    s = _abi.Samples()
    s.model = _abi.MODEL_IDS[model]
    att = samples["attack"]
    s.num_samples, s.num_teams = int(att.shape[0]), int(att.shape[1])
    cs = samples.get("confederation_strength")
    s.num_conferences = 0 if cs is None else int(cs.shape[1])
    src = dict(samples)
    if model in ("dixon_coles", "extended"):
        src["home_attack"] = samples["home_advantage"]  # include/bplx.h: home advantage travels here
    for k in _SAMPLE_KEYS:
        v = src.get(k)
        setattr(s, k, None if v is None else ptr(v))
    return s


def _fixtures_struct(fixtures: dict, ptr):
    f = _abi.Fixtures()
    f.num_fixtures = int(len(fixtures["home_team"]))
    for k, _ in _FIXTURE_KEYS:
        v = fixtures.get(k)
        setattr(f, k, None if v is None else ptr(v))
    return f


def score_grid_workspace_bytes(model: str, samples: Dict[str, torch.Tensor], fixtures: Dict[str, torch.Tensor],
                               max_goals: int) -> int:
    """Bytes of device workspace ``score_grid`` needs for these shapes (``bplx_score_grid_workspace_bytes``)."""
    lib = _abi.lib()
    s = _samples_struct(model, samples, lambda t: int(t.data_ptr()))
    f = _fixtures_struct(fixtures, lambda t: int(t.data_ptr()))
    return int(lib.bplx_score_grid_workspace_bytes(C.byref(s), C.byref(f), max_goals))


def score_grid(model: str, samples: Dict[str, torch.Tensor], fixtures: Dict[str, torch.Tensor], max_goals: int,
               scale: Optional[float] = None, want_outcome: bool = True, workspace=None, grid=None, outcome=None,
               stream=None, reuse_tables: bool = False):
    """Device path: posterior arrays ``[S, T]`` float32 and fixture index tensors on the GPU.
    Returns ``(grid [F, g, g], outcome [F, 3] or None)``; ``scale`` defaults to ``1/S``.  ``reuse_tables``: the
    ``workspace`` passed in still holds the exponential tables a previous call built from these very samples."""
    lib = _abi.lib()
    keep = []

    def ptr(t):
        assert t.is_cuda and t.is_contiguous()
        keep.append(t)
        return int(t.data_ptr())

    for k in _SAMPLE_KEYS + ("home_advantage",):
        if samples.get(k) is not None:
            assert samples[k].dtype == torch.float32
    for k, dt in _FIXTURE_KEYS:
        if fixtures.get(k) is not None:
            assert fixtures[k].dtype == (torch.uint16 if dt == np.uint16 else torch.uint8), k
    s = _samples_struct(model, samples, ptr)
    f = _fixtures_struct(fixtures, ptr)
    g = max_goals + 1
    dev = samples["attack"].device
    need = int(lib.bplx_score_grid_workspace_bytes(C.byref(s), C.byref(f), max_goals))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
    grid = torch.empty((f.num_fixtures, g, g), dtype=torch.float32, device=dev) if grid is None else grid
    if want_outcome and outcome is None:
        outcome = torch.empty((f.num_fixtures, 3), dtype=torch.float32, device=dev)
    scale = 1.0 / s.num_samples if scale is None else scale
    _abi.check(lib.bplx_score_grid_ex(C.byref(s), C.byref(f), max_goals, C.c_float(scale), _dptr(grid),
                                      _dptr(outcome) if want_outcome else None, _dptr(workspace), workspace.numel(),
                                      _stream_ptr(stream), 1 if reuse_tables else 0))
    return grid, (outcome if want_outcome else None)


def score_grid_host(model: str, samples: Dict[str, np.ndarray], fixtures: Dict[str, np.ndarray], max_goals: int,
                    scale: Optional[float] = None, want_outcome: bool = True):
    """Host path: numpy in, numpy out (copies inside the call)."""
    lib = _abi.lib()
    keep = []

    def ptr(a):
        keep.append(a)
        return a.ctypes.data

    smp = {k: (None if v is None else np.ascontiguousarray(v, dtype=np.float32)) for k, v in samples.items()}
    fx = {k: (None if fixtures.get(k) is None else np.ascontiguousarray(fixtures[k], dtype=dt)) for k, dt in _FIXTURE_KEYS}
    s = _samples_struct(model, smp, ptr)
    f = _fixtures_struct(fx, ptr)
    g = max_goals + 1
    grid = np.empty((f.num_fixtures, g, g), dtype=np.float32)
    outcome = np.empty((f.num_fixtures, 3), dtype=np.float32) if want_outcome else None
    scale = 1.0 / s.num_samples if scale is None else scale
    _abi.check(lib.bplx_score_grid_host(C.byref(s), C.byref(f), max_goals, C.c_float(scale), grid.ctypes.data,
                                        None if outcome is None else outcome.ctypes.data))
    return grid, outcome
