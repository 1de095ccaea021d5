"""Synthetic inputs: the reference's test fixtures and the BASELINE.json configs.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``dummy_data`` / ``timed_dummy_data`` / ``neutral_dummy_data`` regenerate the data of
``/root/reference/tests/conftest.py:7-29,32-62,65-116`` with the same seeds and draw order.
``config_N`` build the synthetic leagues of SURVEY.md section 8(d).
Returned dicts use the reference's ``training_data`` keys.
"""
from __future__ import annotations

import itertools

import numpy as np


def dummy_data():
    """``tests/conftest.py:7-29``: 20 teams, all 380 ordered pairs, Poisson(2.1)/(1.7), seed 42."""
    np.random.seed(42)
    home_goals = np.random.poisson(2.1, size=380)
    away_goals = np.random.poisson(1.7, size=380)
    teams = [str(i) for i in range(20)]
    home_team, away_team = [], []
    for a, b in itertools.permutations(teams, 2):
        home_team.append(a)
        away_team.append(b)
    return {"home_team": home_team, "away_team": away_team, "home_goals": home_goals, "away_goals": away_goals}


def timed_dummy_data():
    """``tests/conftest.py:32-62``: two teams, 60 matches, A wins / draws / B wins phases."""
    mpp = 20
    home_team = ["A", "B"] * int(mpp / 2) * 3
    away_team = ["B", "A"] * int(mpp / 2) * 3
    home_goals = [2, 0] * int(mpp / 2) + [1] * mpp + [0, 2] * int(mpp / 2)
    away_goals = [0, 2] * int(mpp / 2) + [1] * mpp + [2, 0] * int(mpp / 2)
    time_diff = np.linspace(5, 0, num=mpp * 3)
    return {"home_team": home_team, "away_team": away_team, "home_goals": home_goals,
            "away_goals": away_goals, "time_diff": time_diff}


def neutral_dummy_data():
    """``tests/conftest.py:65-116``: 380 league + 190 neutral cup matches, seed 42."""
    np.random.seed(42)
    neutral_venue = np.array([0] * 380 + [1] * 190)
    home_means = [2.1 if v == 0 else 1.9 for v in neutral_venue]
    away_means = [1.7 if v == 0 else 1.9 for v in neutral_venue]
    home_goals = np.random.poisson(home_means)
    away_goals = np.random.poisson(away_means)
    time_diff = np.concatenate([np.array([1.0] * 380), np.linspace(0, 10, num=190)])
    game_weights = np.concatenate([np.array([1.0] * 380), np.random.uniform(0, 10, size=190)])
    teams = [str(i) for i in range(20)]
    home_team, away_team = [], []
    for a, b in itertools.permutations(teams, 2):
        home_team.append(a)
        away_team.append(b)
    for a, b in itertools.combinations(teams, 2):
        home_team.append(a)
        away_team.append(b)
    home_conf = [str(int(ht) // 4) for ht in home_team]
    away_conf = [str(int(at) // 4) for at in away_team]
    return {"home_team": home_team, "away_team": away_team, "home_conf": home_conf, "away_conf": away_conf,
            "home_goals": home_goals, "away_goals": away_goals, "neutral_venue": neutral_venue,
            "time_diff": time_diff, "game_weights": game_weights}


def with_covariates(td, K=3, seed=5):
    """A reference-format data set plus `team_covariates` (dict team -> K features of different scales and offsets)."""
    td = dict(td)
    teams = sorted(set(td["home_team"]) | set(td["away_team"]))
    rng = np.random.default_rng(seed)
    td["team_covariates"] = {t: rng.normal(size=K) * (1.0 + np.arange(K)) + np.arange(K) for t in teams}
    return td


def ref_shim_cases():
    """name -> (model, data set, fit kwargs): the cases of scripts/make_ref_shim_golden.py / tests/test_golden.py."""
    return {
        "dixon_coles": ("dixon_coles", dummy_data(), {}),
        "extended": ("extended", dummy_data(), {}),
        "extended_weighted_cov": ("extended", with_covariates(timed_dummy_data()), dict(epsilon=0.3, rescale_weights=True)),
        "neutral": ("neutral", neutral_dummy_data(), {}),
        "neutral_weighted": ("neutral", neutral_dummy_data(), dict(epsilon=0.2, rescale_weights=True)),
        "neutral_wc": ("neutral_wc", neutral_dummy_data(), dict(epsilon=0.2)),
        # the data of BASELINE.json's configs[1] and configs[2], as bench.py prepares them
        "config_2": ("extended", config_2(), dict(epsilon=0.01)),
        "config_3": ("neutral_wc", config_3(), dict(epsilon=0.1)),
    }


def _names(n):
    w = len(str(n - 1))
    return [f"T{str(i).zfill(w)}" for i in range(n)]


def config_2(seed=1002, seasons=5, T=20, K=3):
    """Extended: T=20, 5 seasons x 380, K=3 covariates, time decay (SURVEY 8d cfg 2)."""
    rng = np.random.default_rng(seed)
    names = _names(T)
    att = rng.normal(0, 0.3, T); dfn = rng.normal(0, 0.3, T); gam = rng.normal(0.25, 0.1, T)
    X = rng.normal(0, 1, (T, K))
    ht, at, hg, ag, td = [], [], [], [], []
    for s in range(seasons):
        pairs = list(itertools.permutations(range(T), 2))
        weeks = np.linspace(38, 0, len(pairs))
        for (h, a), wk in zip(pairs, weeks):
            ht.append(names[h]); at.append(names[a])
            hg.append(rng.poisson(np.exp(att[h] - dfn[a] + gam[h])))
            ag.append(rng.poisson(np.exp(att[a] - dfn[h])))
            td.append(52.0 * (seasons - 1 - s) + wk)
    return {"home_team": ht, "away_team": at, "home_goals": np.array(hg), "away_goals": np.array(ag),
            "time_diff": np.array(td), "team_covariates": {names[i]: X[i] for i in range(T)}}


def config_3(seed=1003, T=220, M=40000, Cf=6):
    """NeutralWC: T=220, M=40k random pairs, conf = team mod 6, time/game weights (cfg 3)."""
    rng = np.random.default_rng(seed)
    names = _names(T)
    att = rng.normal(0, 0.3, T); dfn = rng.normal(0, 0.3, T)
    ha = rng.normal(0.1, 0.1, T); aa = rng.normal(-0.1, 0.1, T)
    hd = rng.normal(0.1, 0.1, T); ad = rng.normal(-0.1, 0.1, T)
    conf = rng.normal(0, 0.2, Cf)
    h = rng.integers(0, T, M)
    a = (h + rng.integers(1, T, M)) % T
    nv = (rng.random(M) < 0.35).astype(np.int64)
    n = 1 - nv
    eh = att[h] - dfn[a] + conf[h % Cf] - conf[a % Cf] + n * ha[h] - n * ad[a]
    ea = att[a] - dfn[h] + conf[a % Cf] - conf[h % Cf] + n * aa[a] - n * hd[h]
    hg = rng.poisson(np.exp(eh)); ag = rng.poisson(np.exp(ea))
    return {"home_team": [names[i] for i in h], "away_team": [names[i] for i in a],
            "home_conf": [f"C{i % Cf}" for i in h], "away_conf": [f"C{i % Cf}" for i in a],
            "home_goals": hg, "away_goals": ag, "neutral_venue": nv,
            "time_diff": rng.uniform(0, 20, M), "game_weights": rng.integers(1, 5, M).astype(np.float64)}


def config_4(seed=1004, T=20, G=30):
    """Dynamic: 30 seasons x 380, gameweek = season index, random-walk strengths (cfg 4)."""
    rng = np.random.default_rng(seed)
    names = _names(T)
    att = rng.normal(0, 0.3, T); dfn = rng.normal(0, 0.3, T)
    ht, at, hg, ag, gw = [], [], [], [], []
    for g in range(G):
        if g:
            att = att + rng.normal(0, 0.1, T); dfn = dfn + rng.normal(0, 0.1, T)
        for h, a in itertools.permutations(range(T), 2):
            ht.append(names[h]); at.append(names[a]); gw.append(g)
            hg.append(rng.poisson(np.exp(att[h] - dfn[a] + 0.2)))
            ag.append(rng.poisson(np.exp(att[a] - dfn[h] - 0.2)))
    return {"home_team": ht, "away_team": at, "home_goals": np.array(hg), "away_goals": np.array(ag),
            "gameweek": np.array(gw), "neutral_venue": np.zeros(len(ht), dtype=np.int64)}


def config_5(seed=1005, T=220, Cf=6, S=16384, F=10000):
    """Predict grid: S posterior samples for a WC model + F random fixtures (cfg 5).

    Returns (samples dict of float32 arrays, fixtures dict of index arrays).
    """
    rng = np.random.default_rng(seed)
    f32 = np.float32
    samples = {
        "attack": rng.normal(0, 0.3, (S, T)).astype(f32), "defence": rng.normal(0, 0.3, (S, T)).astype(f32),
        "home_attack": rng.normal(0.1, 0.1, (S, T)).astype(f32), "away_attack": rng.normal(-0.1, 0.1, (S, T)).astype(f32),
        "home_defence": rng.normal(0.1, 0.1, (S, T)).astype(f32), "away_defence": rng.normal(-0.1, 0.1, (S, T)).astype(f32),
        "confederation_strength": rng.normal(0, 0.2, (S, Cf)).astype(f32),
        "corr_coef": rng.uniform(-0.1, 0.1, S).astype(f32),
    }
    h = rng.integers(0, T, F)
    a = (h + rng.integers(1, T, F)) % T
    fixtures = {"home_team": h.astype(np.uint16), "away_team": a.astype(np.uint16),
                "home_conf": (h % Cf).astype(np.uint8), "away_conf": (a % Cf).astype(np.uint8),
                "neutral_venue": (rng.random(F) < 0.35).astype(np.uint8)}
    return samples, fixtures
