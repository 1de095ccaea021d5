"""Host-side data prep: what each reference ``fit`` does before handing arrays to ``_model``.

Mirrors ``parse_teams`` (``bpl/_util.py:115-135``), the conference lookup of
``neutral_dixon_coles_WC.py:253-265``, the covariate checks (``extended_dixon_coles.py:282-289``)
and the weight formulas of the four ``_model``s (``extended_dixon_coles.py:202-205``,
``neutral_dixon_coles.py:251-257``, ``neutral_dixon_coles_WC.py:205-207``).  Index dtypes are the
reference's ``DTYPES`` (``bpl/base.py:16-22``).  O(M) host work, done once per fit.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import numpy as np

from . import _abi

MAX_GOALS = 15  # bpl/base.py:15
DTYPES = {"goals": np.uint8, "teams": np.uint16, "conferences": np.uint8, "venue": np.uint8}


def parse_teams(home_team, away_team):
    """``bpl/_util.py:115-135``: sorted unique names, name->index dict, index arrays."""
    h, a = np.asarray(home_team), np.asarray(away_team)
    if h.dtype.kind in "USiuf" and a.dtype.kind == h.dtype.kind and h.ndim == 1:
        # one sort of the 2M names instead of 2M dictionary look-ups: same sorted unique names, same indices
        teams, inv = np.unique(np.concatenate([h, a]), return_inverse=True)
        home_ind, away_ind = inv[:len(h)], inv[len(h):]
    else:  # names numpy cannot order as one array (mixed types): the reference's own loop
        teams = np.array(sorted(set(home_team) | set(away_team)))
        lut = {t: i for i, t in enumerate(teams)}
        home_ind = np.array([lut[t] for t in home_team])
        away_ind = np.array([lut[t] for t in away_team])
    if len(teams) > np.iinfo(DTYPES["teams"]).max:
        raise ValueError("too many teams for the uint16 team index")
    teams_dict = {t: i for i, t in enumerate(teams)}
    return teams, teams_dict, home_ind.astype(DTYPES["teams"]), away_ind.astype(DTYPES["teams"])


def _goals(x):
    a = np.asarray(x)
    if a.size and (a.min() < 0 or a.max() > 255):
        raise ValueError("goals must be in [0, 255]")
    return a.astype(DTYPES["goals"])


@dataclass
class MatchArrays:
    """The positional arguments of a reference ``_model`` after host prep."""

    model: str
    num_teams: int
    home_team: np.ndarray
    away_team: np.ndarray
    home_goals: np.ndarray
    away_goals: np.ndarray
    weights: Optional[np.ndarray] = None        # float32 [M]; None = unweighted branch
    neutral_venue: Optional[np.ndarray] = None  # uint8 [M]
    home_conf: Optional[np.ndarray] = None
    away_conf: Optional[np.ndarray] = None
    num_conferences: int = 0
    covariates: Optional[np.ndarray] = None     # float32 [T, K], already standardised
    gameweek: Optional[np.ndarray] = None
    num_gameweeks: int = 0
    as_written: bool = False                    # dynamic only: BPLX_FLAG_DYNAMIC_AS_WRITTEN
    _keep: list = field(default_factory=list, repr=False)

    @property
    def num_matches(self) -> int:
        return int(len(self.home_team))

    @property
    def num_covariates(self) -> int:
        return 0 if self.covariates is None else int(self.covariates.shape[1])

    def desc(self) -> _abi.ProblemDesc:
        """ctypes ``bplx_problem_desc`` pointing at this object's arrays (kept alive by it)."""
        d = _abi.ProblemDesc()
        d.model = _abi.MODEL_IDS[self.model]
        d.num_matches = self.num_matches
        d.num_teams = self.num_teams
        d.num_covariates = self.num_covariates
        d.num_conferences = self.num_conferences
        d.num_gameweeks = self.num_gameweeks
        d.flags = 1 if (self.model == "dynamic" and self.as_written) else 0

        def ptr(a, dtype, ctype):
            if a is None:
                return None
            arr = np.ascontiguousarray(a, dtype=dtype)
            self._keep.append(arr)
            return arr.ctypes.data_as(C.POINTER(ctype))

        d.home_team = ptr(self.home_team, np.uint16, C.c_uint16)
        d.away_team = ptr(self.away_team, np.uint16, C.c_uint16)
        d.home_goals = ptr(self.home_goals, np.uint8, C.c_uint8)
        d.away_goals = ptr(self.away_goals, np.uint8, C.c_uint8)
        d.neutral_venue = ptr(self.neutral_venue, np.uint8, C.c_uint8)
        d.home_conf = ptr(self.home_conf, np.uint8, C.c_uint8)
        d.away_conf = ptr(self.away_conf, np.uint8, C.c_uint8)
        d.gameweek = ptr(self.gameweek, np.int32, C.c_int32)
        d.weights = ptr(self.weights, np.float32, C.c_float)
        d.covariates = ptr(self.covariates, np.float32, C.c_float)
        return d


def _covariates(team_covariates, teams):
    """``extended_dixon_coles.py:282-289`` + the standardisation of ``:124-127`` (float32)."""
    if not team_covariates:
        return None, None, None
    if set(team_covariates.keys()) != set(teams):
        raise ValueError("team_covariates must contain all the teams in the data.")
    X = np.array([team_covariates[t] for t in teams], dtype=np.float32)
    mean, std = X.mean(axis=0), X.std(axis=0)
    return ((X - mean) / std).astype(np.float32), mean, std


def prepare(model: str, training_data: Dict[str, Any], epsilon=None, rescale_weights: bool = False):
    """Returns (MatchArrays, meta) where meta holds teams / dict / conferences / covariate stats."""
    teams, teams_dict, home_ind, away_ind = parse_teams(training_data["home_team"], training_data["away_team"])
    hg, ag = _goals(training_data["home_goals"]), _goals(training_data["away_goals"])
    M = len(home_ind)
    meta: Dict[str, Any] = {"teams": teams, "teams_dict": teams_dict}
    arr = MatchArrays(model=model, num_teams=len(teams), home_team=home_ind, away_team=away_ind,
                      home_goals=hg, away_goals=ag)
    if model == "dixon_coles":
        return arr, meta
    Xs, xm, xs = _covariates(training_data.get("team_covariates"), teams)
    arr.covariates = Xs
    meta["team_covariates_mean"], meta["team_covariates_std"] = xm, xs
    if model == "extended":
        if epsilon is not None:  # extended_dixon_coles.py:202-205
            td = np.asarray(training_data["time_diff"], dtype=np.float32)
            w = np.exp(-np.float32(epsilon) * td)
            if rescale_weights:
                w = M * w / w.sum()
            arr.weights = w.astype(np.float32)
        return arr, meta
    if model in ("neutral", "neutral_wc"):
        arr.neutral_venue = np.asarray(training_data["neutral_venue"]).astype(DTYPES["venue"])
        gw = np.asarray(training_data["game_weights"], dtype=np.float32)
        if model == "neutral":  # neutral_dixon_coles.py:308-318: read with .get(), required only when epsilon is set
            td = training_data.get("time_diff")
            if td is None:
                if epsilon is not None:
                    raise ValueError("time_diff must be provided in training_data to include exponential time decay in model.")
                td = np.zeros(M, dtype=np.float32)
        else:  # neutral_dixon_coles_WC.py:270: read unconditionally
            td = training_data["time_diff"]
        td = np.asarray(td, dtype=np.float32)
        if model == "neutral":  # neutral_dixon_coles.py:251-257
            w = np.ones(M, dtype=np.float32)
            if epsilon is not None:
                w = w * np.exp(-np.float32(epsilon) * td)
                if rescale_weights:
                    w = M * w / w.sum()
            w = w * gw
        else:  # neutral_dixon_coles_WC.py:205-207 (epsilon is a float, default 0.0)
            w = np.exp(-np.float32(epsilon if epsilon is not None else 0.0) * td) * gw
            if rescale_weights:
                w = M * w / w.sum()
            confs = np.array(sorted(set(training_data["home_conf"]) | set(training_data["away_conf"])))
            cdict = {c: i for i, c in enumerate(confs)}
            arr.home_conf = np.array([cdict[c] for c in training_data["home_conf"]], DTYPES["conferences"])
            arr.away_conf = np.array([cdict[c] for c in training_data["away_conf"]], DTYPES["conferences"])
            arr.num_conferences = len(confs)
            meta["conferences"], meta["conferences_dict"] = confs, cdict
        arr.weights = w.astype(np.float32)
        return arr, meta
    if model == "dynamic":  # dynamic_dixon_coles.py:262-296 (gameweeks 0-based, G = max + 1: SURVEY D2)
        arr.neutral_venue = np.asarray(training_data["neutral_venue"]).astype(DTYPES["venue"])
        gw = np.asarray(training_data["gameweek"], dtype=np.int32)
        if gw.size and gw.min() < 0:
            raise ValueError("gameweek must be >= 0")
        arr.gameweek = gw
        arr.num_gameweeks = int(gw.max()) + 1
        return arr, meta
    raise ValueError(f"unknown model {model!r}")
