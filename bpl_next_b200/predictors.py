"""The reference's predictor classes on top of the CUDA path: same names, ``fit`` / ``predict_*`` signatures and
fitted attribute names as ``bpl/dixon_coles.py``, ``bpl/extended_dixon_coles.py``, ``bpl/neutral_dixon_coles.py`` and
``bpl/neutral_dixon_coles_WC.py`` (+ ``bpl/base.py``), so the reference's tests read the same against this package.

``fit`` = host prep (``data.prepare``) -> ``Problem`` (K1) -> batched GPU NUTS (``nuts.sample``, standing in for
``numpyro.infer.NUTS`` / ``MCMC``) -> constrained samples cut by the library's site layout -> the deterministic sites the
reference records (``attack``, ``defence``, venue effects, ``rho``, ``corr_coef``).  ``predict_*`` = K3.
``mcmc_kwargs`` understands ``num_chains`` (default 1 like numpyro; use hundreds on a GPU) and ``thin``.
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, Optional, Union

import warnings
from datetime import datetime

import numpy as np
import torch

from . import data as bdata
from . import nuts as bnuts
from . import parallel
from .problem import Problem, score_grid_host

MAX_GOALS = bdata.MAX_GOALS


def _str_to_list(*args):
    return tuple([a] if isinstance(a, (str, int, np.integer)) else list(a) for a in args)  # bpl/_util.py:10-14


def constrain(flat: np.ndarray, layout: Dict[str, tuple]) -> Dict[str, np.ndarray]:
    """Flat unconstrained draws ``[S, D]`` -> dict of constrained sites (numpyro ``biject_to``: exp / clipped sigmoid)."""
    out = {}
    fi = np.finfo(np.float32)
    for name, (off, cnt, tr) in layout.items():
        x = flat[:, off:off + cnt].astype(np.float32)
        if tr == "exp":
            x = np.exp(x)
        elif tr == "sigmoid":
            x = np.clip(1.0 / (1.0 + np.exp(-x)), fi.tiny, 1.0 - fi.eps).astype(np.float32)
        out[name] = x[:, 0] if cnt == 1 and not name.endswith(("_decentered", "coefficients")) and name not in (
            "standardised_attack", "standardised_defence") else x
    return out


def _run_chains(p: Problem, num_chains: int, random_state: int, num_warmup: int, num_samples: int, thin: int, kw: dict,
                owner) -> tuple:
    """NUTS on ``num_chains`` chains of problem ``p``.  Under ``torch.distributed`` (one process per GPU) the chains are
    partitioned over the ranks -- no collective while sampling -- and the draws are all-gathered at the end, so every
    rank ends with the same ``[num_chains * num_keep, D]`` array (chains concatenated like ``MCMC.get_samples()``)."""
    rank, nranks = parallel.world()
    c0, cn = parallel.shard(num_chains, rank, nranks)
    if cn == 0:
        raise ValueError(f"num_chains={num_chains} is smaller than the number of ranks ({nranks})")
    g = torch.Generator(device="cuda").manual_seed(int(random_state))
    theta0 = torch.rand((p.D, num_chains), generator=g, device="cuda") * 4.0 - 2.0  # numpyro init_to_uniform(radius=2)
    theta0 = theta0[:, c0:c0 + cn].contiguous()

    def potential(theta, lp, grad):
        p.logdensity(theta, chain_minor=True, lp=lp, grad=grad)

    def potential_cm(theta, lp, grad):  # [chains, D] tensors: large models keep the sampler state chain-major
        p.logdensity(theta, chain_minor=False, lp=lp, grad=grad)

    kw.setdefault("potential_cm", potential_cm)
    run = bnuts.sample(potential, theta0, num_warmup=num_warmup, num_samples=num_samples, thin=thin,
                       seed=int(random_state), chain_offset=c0, **kw)
    owner.nuts_run = run
    K, D, C = run.samples.shape
    flat_dev = run.samples.permute(2, 0, 1).reshape(C * K, D).contiguous()
    if nranks > 1:
        import torch.distributed as dist
        sizes = [parallel.shard(num_chains, r, nranks)[1] * K for r in range(nranks)]
        parts = [torch.empty((n, D), dtype=flat_dev.dtype, device=flat_dev.device) for n in sizes]
        dist.all_gather(parts, flat_dev)  # the final sample gather (NCCL)
        flat_dev = torch.cat(parts, dim=0)
    _, _, cc = p.logdensity(flat_dev)  # the deterministic site "corr_coef" of every draw
    torch.cuda.synchronize()
    return flat_dev, cc


class _BplxPredictor:
    model = ""

    def __init__(self):
        self.teams = None
        self._teams_dict = None
        self.problem: Optional[Problem] = None
        self.nuts_run = None

    # ---- fit ----------------------------------------------------------------------------------------------
    def _fit(self, training_data, random_state, num_warmup, num_samples, mcmc_kwargs, epsilon=None,
             rescale_weights=False):
        kw = dict(mcmc_kwargs or {})
        num_chains = int(kw.pop("num_chains", 1))
        thin = int(kw.pop("thin", kw.pop("thinning", 1)))
        kw.pop("chain_method", None)
        kw.pop("progress_bar", None)
        arr, meta = bdata.prepare(self.model, training_data, epsilon=epsilon, rescale_weights=rescale_weights)
        self.teams, self._teams_dict = meta["teams"], meta["teams_dict"]
        self._meta = meta
        self.problem = p = Problem(arr)
        flat_dev, cc = _run_chains(p, num_chains, random_state, num_warmup, num_samples, thin, kw, self)
        s = constrain(flat_dev.cpu().numpy(), p.layout)
        s["corr_coef"] = cc.cpu().numpy()
        return arr, s

    def fit_streaming(self, training_data, epsilon=None, rescale_weights: bool = False, random_state: int = 42,
                      num_warmup: int = 500, num_samples: int = 1000, num_chains: int = 1024, thin: int = 10,
                      diag_lags: int = 24, max_tree_depth: int = 10, max_launches: Optional[int] = None,
                      set_posterior: bool = False, state_layout: str = "auto") -> Dict[str, Any]:
        """Many-chain ``fit`` that never stores all draws: the chains are partitioned over the ``torch.distributed`` ranks
        (no collective while sampling); every chain keeps streaming moment / lagged-product accumulators in the step
        kernel and only every ``thin``-th draw is stored.  At the end the R-hat / ESS moments are all-reduced
        (``diagnostics.streaming_summary``) and the per-chain sampler summaries all-gathered; with ``set_posterior`` the
        thinned draws are gathered too and become the fitted attributes (the ``[S, T]`` layout of
        ``neutral_dixon_coles_WC.py:308-334``), so ``predict_*`` works as after ``fit``.  Returns the run summary."""
        import time as _time

        from . import diagnostics as dg

        if self.model == "dixon_coles":
            arr, meta = bdata.prepare(self.model, training_data)
        else:
            arr, meta = bdata.prepare(self.model, training_data, epsilon=epsilon, rescale_weights=rescale_weights)
        self.teams, self._teams_dict, self._meta = meta["teams"], meta["teams_dict"], meta
        self.problem = p = Problem(arr)
        rank, nranks = parallel.world()
        c0, cn = parallel.shard(num_chains, rank, nranks)
        if cn == 0:
            raise ValueError(f"num_chains={num_chains} is smaller than the number of ranks ({nranks})")
        # numpyro init_to_uniform(radius=2), a pure function of (seed, global chain index range)
        g = torch.Generator(device="cuda").manual_seed(int(random_state) + 7919 * rank)
        theta0 = (torch.rand((p.D, cn), generator=g, device="cuda") * 4.0 - 2.0).contiguous()

        def potential(theta, lp, grad):
            p.logdensity(theta, chain_minor=True, lp=lp, grad=grad)

        def potential_cm(theta, lp, grad):  # [chains, D] tensors: large models keep the sampler state chain-major
            p.logdensity(theta, chain_minor=False, lp=lp, grad=grad)

        torch.cuda.synchronize()
        t0 = _time.perf_counter()
        run = bnuts.sample(potential, theta0, num_warmup=num_warmup, num_samples=num_samples, thin=thin,
                           seed=int(random_state), chain_offset=c0, max_tree_depth=max_tree_depth,
                           max_launches=max_launches, diag_lags=diag_lags, potential_cm=potential_cm,
                           state_layout=state_layout)
        torch.cuda.synchronize()
        t_sample = _time.perf_counter() - t0
        self.nuts_run = run
        complete_local = bool((run.transitions >= num_warmup + num_samples).all())
        # ---- the exchange: diagnostics moments (all-reduce), per-chain sampler summaries (all-gather) ---------------
        t0 = _time.perf_counter()
        summ = dg.streaming_summary(run.diag)
        per_chain = torch.from_numpy(np.stack([run.step_size, run.num_divergent.astype(np.float32),
                                               run.num_leapfrog.astype(np.float32),
                                               run.transitions.astype(np.float32)], axis=1).astype(np.float32)).cuda()
        if nranks > 1:
            import torch.distributed as dist
            parts = [torch.empty((parallel.shard(num_chains, r, nranks)[1], 4), dtype=torch.float32, device="cuda")
                     for r in range(nranks)]
            dist.all_gather(parts, per_chain)
            per_chain = torch.cat(parts, dim=0)
            launches = torch.tensor([float(run.launches)], device="cuda")
            dist.all_reduce(launches, op=dist.ReduceOp.MAX)
            launches = int(launches.item())
        else:
            launches = int(run.launches)
        torch.cuda.synchronize()
        t_exchange = _time.perf_counter() - t0
        pc = per_chain.cpu().numpy()
        complete = bool((pc[:, 3] >= num_warmup + num_samples).all())
        self.posterior_mean, self.posterior_sd = summ["mean"].cpu().numpy(), summ["sd"].cpu().numpy()
        self.rhat, self.ess = summ.get("rhat"), summ["ess"]
        out = {"complete": complete, "chains": int(pc.shape[0]), "draws_per_chain": num_samples, "thin": thin,
               "stored_draws_per_chain": int(run.samples.shape[0]), "diag_lags": int(summ["lags"]),
               "ess_min": float(summ["ess"].min().item()) if complete else 0.0,
               "ess_median": float(summ["ess"].median().item()) if complete else 0.0,
               "rhat_max": float(summ["rhat"].max().item()) if complete and "rhat" in summ else None,
               "lag_window_hit": int(summ["lag_window_hit"]),
               "leapfrogs_total": float(pc[:, 2].sum()), "leapfrogs_max_chain": float(pc[:, 2].max()),
               "logdensity_launches": launches, "divergences": int(pc[:, 1].sum()),
               "step_size_median": float(np.median(pc[:, 0])),
               "match_evals_total": float(pc[:, 2].sum()) * arr.num_matches,
               "sampling_s": t_sample, "exchange_s": t_exchange,
               "exchange": "all_reduce of (lags + 8) x D moment sums, all_gather of [chains, 4] sampler summaries"
                           + (", all_gather of the thinned draws" if set_posterior else "")}
        if set_posterior:
            K, D, C = run.samples.shape
            flat_dev = run.samples.permute(2, 0, 1).reshape(C * K, D).contiguous()
            if nranks > 1:
                import torch.distributed as dist
                sizes = [parallel.shard(num_chains, r, nranks)[1] * K for r in range(nranks)]
                parts = [torch.empty((n, D), dtype=flat_dev.dtype, device=flat_dev.device) for n in sizes]
                dist.all_gather(parts, flat_dev)
                flat_dev = torch.cat(parts, dim=0)
            _, _, cc = p.logdensity(flat_dev)
            torch.cuda.synchronize()
            s = constrain(flat_dev.cpu().numpy(), p.layout)
            s["corr_coef"] = cc.cpu().numpy()
            self._set_posterior(arr, s)
        return out

    @staticmethod
    def _prior_means(arr, s):
        S = len(s["mean_defence"])
        if arr.covariates is not None:  # extended_dixon_coles.py:138-143
            Xs = arr.covariates.astype(np.float32)
            return s["attack_coefficients"] @ Xs.T, s["mean_defence"][:, None] + s["defence_coefficients"] @ Xs.T
        return np.zeros((S, 1), np.float32), s["mean_defence"][:, None]

    # ---- predict (bpl/base.py:62-348) -------------------------------------------------------------------------
    def _parse_fixture_args(self, home_team, away_team):
        home_team, away_team = _str_to_list(home_team, away_team)
        if isinstance(home_team[0], str):
            home_team = [self._teams_dict[t] for t in home_team]
        if isinstance(away_team[0], str):
            away_team = [self._teams_dict[t] for t in away_team]
        return np.asarray(home_team, dtype=np.uint16), np.asarray(away_team, dtype=np.uint16)

    def _samples(self) -> Dict[str, np.ndarray]:
        raise NotImplementedError

    def _fixture_extras(self, n, **kw) -> Dict[str, np.ndarray]:
        return {}

    def _grid(self, home_team, away_team, max_goals, **kw):
        h, a = self._parse_fixture_args(home_team, away_team)
        fx = {"home_team": h, "away_team": a, **self._fixture_extras(len(h), **kw)}
        grid, outcome = score_grid_host(self.model, self._samples(), fx, max_goals)
        return grid, outcome

    def predict_score_grid_proba(self, home_team, away_team, *args, max_goals: int = MAX_GOALS, **kw):
        grid, _ = self._grid(home_team, away_team, max_goals, **self._extras_from_args(args, kw))
        n = np.arange(0, max_goals + 1)
        hg, ag = np.meshgrid(n, n, indexing="ij")
        return grid, hg, ag

    def predict_outcome_proba(self, home_team, away_team, *args, max_goals: int = MAX_GOALS, knockout: bool = False, **kw):
        _, out = self._grid(home_team, away_team, max_goals, **self._extras_from_args(args, kw))
        hw, dr, aw = out[:, 0], out[:, 1], out[:, 2]
        if knockout:  # neutral_dixon_coles.py:650-653
            norm = hw + aw
            return {"home_win": hw / norm, "away_win": aw / norm}
        return {"home_win": hw, "draw": dr, "away_win": aw}

    def predict_score_proba(self, home_team, away_team, home_goals, away_goals, *args, **kw):
        home_goals, away_goals = _str_to_list(home_goals, away_goals)
        hg, ag = np.asarray(home_goals, dtype=np.int64), np.asarray(away_goals, dtype=np.int64)
        mg = int(max(hg.max(), ag.max(), 1))
        grid, _ = self._grid(home_team, away_team, mg, **self._extras_from_args(args, kw))
        if len(hg) == 1 and grid.shape[0] > 1:
            hg, ag = np.repeat(hg, grid.shape[0]), np.repeat(ag, grid.shape[0])
        return grid[np.arange(grid.shape[0]), hg, ag]

    def _n_proba(self, n, team, opponent, home, max_goals, scored, **kw):
        n = [n] if isinstance(n, (int, np.integer)) else list(np.asarray(n))
        grid, _ = self._grid(team, opponent, max_goals, **kw) if home else self._grid(opponent, team, max_goals, **kw)
        # marginal of the grid: sum over the other side's goals (bpl/base.py:248-348)
        team_axis_is_home = home if scored else not home
        marg = grid.sum(axis=2) if team_axis_is_home else grid.sum(axis=1)  # [F, g]
        return marg[0, np.asarray(n, dtype=np.int64)]

    def predict_score_n_proba(self, n, team, opponent, home: bool = True, max_goals: int = MAX_GOALS, **kw):
        return self._n_proba(n, team, opponent, home, max_goals, True, **kw)

    def predict_concede_n_proba(self, n, team, opponent, home: bool = True, max_goals: int = MAX_GOALS, **kw):
        return self._n_proba(n, team, opponent, home, max_goals, False, **kw)

    def _extras_from_args(self, args, kw):
        return kw

    # ---- simulation from the grid (bpl/base.py:150-246, bpl/_util.py:96-112) -----------------------------------
    @staticmethod
    def _choice(probs: np.ndarray, num_samples: int, random_state) -> np.ndarray:
        """``[F, K]`` probabilities -> ``[F, num_samples]`` category indices, one independent stream per fixture (the
        reference draws with ``jax.random.choice`` under ``vmap``; like it, the row is normalised by its own sum)."""
        if random_state is None:
            random_state = int(datetime.now().timestamp() * 100)
        rng = np.random.default_rng(int(random_state))
        cdf = np.cumsum(probs.astype(np.float64), axis=1)
        u = rng.random((probs.shape[0], int(num_samples))) * cdf[:, -1:]
        idx = (u[:, :, None] >= cdf[:, None, :]).sum(axis=2)
        return np.minimum(idx, probs.shape[1] - 1)

    def _sample_score(self, home_team, away_team, num_samples, random_state, max_goals, **kw):
        grid, hg, ag = _BplxPredictor.predict_score_grid_proba(self, home_team, away_team, max_goals=max_goals, **kw)
        idx = self._choice(grid.reshape(grid.shape[0], -1), num_samples, random_state)
        return {"home_score": hg.ravel()[idx], "away_score": ag.ravel()[idx]}

    def _sample_outcome(self, home_team, away_team, num_samples, random_state, max_goals, knockout=False, **kw):
        h, a = self._parse_fixture_args(home_team, away_team)
        pr = _BplxPredictor.predict_outcome_proba(self, h, a, max_goals=max_goals, knockout=knockout, **kw)
        keys = ("home_win", "away_win") if knockout else ("home_win", "draw", "away_win")
        idx = self._choice(np.stack([pr[k] for k in keys], axis=1), num_samples, random_state)
        names = np.append(np.asarray(self.teams), "Draw")
        draw = len(self.teams)
        hrep, arep = np.repeat(h[:, None], num_samples, axis=1).astype(np.int64), np.repeat(a[:, None], num_samples, axis=1).astype(np.int64)
        away_code = 1 if knockout else 2
        winner = np.where(idx == 0, hrep, np.where(idx == away_code, arep, draw))
        return names[winner]

    def sample_score(self, home_team, away_team, num_samples: int = 1, random_state: Optional[int] = None,
                     max_goals: int = MAX_GOALS):
        """``{"home_score", "away_score"}``, each ``[fixtures, num_samples]`` (``bpl/base.py:150-195``)."""
        return self._sample_score(home_team, away_team, num_samples, random_state, max_goals)

    def sample_outcome(self, home_team, away_team, num_samples: int = 1, random_state: Optional[int] = None,
                       max_goals: int = MAX_GOALS):
        """``[fixtures, num_samples]`` names of the winning team, or ``"Draw"`` (``bpl/base.py:197-246``)."""
        return self._sample_outcome(home_team, away_team, num_samples, random_state, max_goals)

    # ---- a team the model has not seen (extended_dixon_coles.py:401-462 and the neutral siblings) ---------------
    def _new_team_pair(self, team_name, team_covariates):
        """Draws of (attack, defence) for a new team from the fitted hierarchical prior, one per posterior sample."""
        if team_name in self.teams:
            raise ValueError(f"Team {team_name} already known to model.")
        if self.attack_coefficients is not None:
            if team_covariates is None:
                warnings.warn(f"You haven't provided features for {team_name}. Assuming team_covariates are the average of "
                              "known teams. For better forecasts, provide team_covariates.")
                x = np.zeros(self.attack_coefficients.shape[1], np.float32)
            else:
                x = (0.5 * (np.asarray(team_covariates, np.float32) - self._meta["team_covariates_mean"])
                     / self._meta["team_covariates_std"]).ravel()
            mean_attack = self.attack_coefficients @ x
            mean_defence = self.mean_defence + self.defence_coefficients @ x
        else:
            mean_attack, mean_defence = 0.0, self.mean_defence
        a_tilde = np.random.normal(loc=0.0, scale=1.0, size=len(self.std_attack))
        b_tilde = np.random.normal(loc=self.rho * a_tilde, scale=np.sqrt(1.0 - self.rho ** 2.0))
        return mean_attack + a_tilde * self.std_attack, mean_defence + b_tilde * self.std_defence

    def _append_team(self, team_name, columns: Dict[str, np.ndarray]):
        self.teams = np.append(self.teams, team_name)
        self._teams_dict[team_name] = len(self._teams_dict)
        for name, col in columns.items():
            setattr(self, name, np.concatenate((getattr(self, name), np.asarray(col, np.float32)[:, None]), axis=1))


class DixonColesMatchPredictor(_BplxPredictor):
    """``bpl/dixon_coles.py:26-163``."""
    model = "dixon_coles"

    def fit(self, training_data, random_state: int = 42, num_warmup: int = 500, num_samples: int = 1000,
            mcmc_kwargs: Optional[Dict[str, Any]] = None, run_kwargs: Optional[Dict[str, Any]] = None):
        arr, s = self._fit(training_data, random_state, num_warmup, num_samples, mcmc_kwargs)
        return self._set_posterior(arr, s)

    def _set_posterior(self, arr, s):
        """Constrained draws of the latent sites -> the sites the reference records (dixon_coles.py:52-61)."""
        self.attack = s["std_attack"][:, None] * s["attack_decentered"]
        self.defence = s["mean_defence"][:, None] + s["std_defence"][:, None] * s["defence_decentered"]
        self.home_advantage = s["home_advantage"]
        self.corr_coef = s["corr_coef"]
        return self

    def _samples(self):
        return {"attack": self.attack, "defence": self.defence, "home_advantage": self.home_advantage,
                "corr_coef": self.corr_coef}


class ExtendedDixonColesMatchPredictor(_BplxPredictor):
    """``bpl/extended_dixon_coles.py:28-462``."""
    model = "extended"

    def fit(self, training_data, random_state: int = 42, num_warmup: int = 500, num_samples: int = 1000,
            epsilon: Optional[float] = None, rescale_weights: bool = False,
            mcmc_kwargs: Optional[Dict[str, Any]] = None, run_kwargs: Optional[Dict[str, Any]] = None):
        arr, s = self._fit(training_data, random_state, num_warmup, num_samples, mcmc_kwargs, epsilon, rescale_weights)
        return self._set_posterior(arr, s)

    def _set_posterior(self, arr, s):
        """Constrained draws of the latent sites -> the sites the reference records (extended_dixon_coles.py:182-187)."""
        am, dm = self._prior_means(arr, s)
        self.attack = am + s["standardised_attack"] * s["std_attack"][:, None]
        self.defence = dm + s["standardised_defence"] * s["std_defence"][:, None]
        self.home_advantage = s["mean_home_advantage"][:, None] + s["std_home_advantage"][:, None] * s["home_advantage_decentered"]
        self.corr_coef = s["corr_coef"]
        self.u = s["u"]
        self.rho = 2.0 * s["u"] - 1.0
        self.mean_defence, self.std_attack, self.std_defence = s["mean_defence"], s["std_attack"], s["std_defence"]
        self.mean_home_advantage, self.std_home_advantage = s["mean_home_advantage"], s["std_home_advantage"]
        self.standardised_attack, self.standardised_defence = s["standardised_attack"], s["standardised_defence"]
        self.attack_coefficients = s.get("attack_coefficients") if arr.covariates is not None else None
        self.defence_coefficients = s.get("defence_coefficients") if arr.covariates is not None else None
        return self

    def _samples(self):
        return {"attack": self.attack, "defence": self.defence, "home_advantage": self.home_advantage,
                "corr_coef": self.corr_coef}

    def add_new_team(self, team_name: str, team_covariates: Optional[np.ndarray] = None) -> None:
        """Parameters for a team not seen in training, drawn from the fitted priors (``extended_dixon_coles.py:401-462``;
        like the reference this uses numpy's global random state)."""
        attack, defence = self._new_team_pair(team_name, team_covariates)
        home_advantage = np.random.normal(loc=self.mean_home_advantage, scale=self.std_home_advantage)
        self._append_team(team_name, {"attack": attack, "defence": defence, "home_advantage": home_advantage})


class NeutralDixonColesMatchPredictor(_BplxPredictor):
    """``bpl/neutral_dixon_coles.py:31-902`` (signatures as there: ``neutral_venue`` is a positional fixture argument)."""
    model = "neutral"
    _effects = ("home_attack", "away_attack", "home_defence", "away_defence")

    def fit(self, training_data, epsilon: Optional[float] = None, rescale_weights: bool = False, random_state: int = 42,
            num_warmup: int = 500, num_samples: int = 1000, mcmc_kwargs: Optional[Dict[str, Any]] = None,
            run_kwargs: Optional[Dict[str, Any]] = None):
        arr, s = self._fit(training_data, random_state, num_warmup, num_samples, mcmc_kwargs, epsilon, rescale_weights)
        return self._set_posterior(arr, s)

    def _set_posterior(self, arr, s):
        """Constrained draws of the latent sites -> the sites the reference records."""
        am, dm = self._prior_means(arr, s)
        self.attack = am + s["standardised_attack"] * s["std_attack"][:, None]
        self.defence = dm + s["standardised_defence"] * s["std_defence"][:, None]
        for nm in self._effects:  # neutral_dixon_coles.py:205-224
            setattr(self, nm, s["mean_" + nm][:, None] + s["std_" + nm][:, None] * s[nm + "_decentered"])
            setattr(self, "mean_" + nm, s["mean_" + nm])
            setattr(self, "std_" + nm, s["std_" + nm])
        self.corr_coef = s["corr_coef"]
        self.u = s["u"]
        self.rho = 2.0 * s["u"] - 1.0
        self.mean_defence, self.std_attack, self.std_defence = s["mean_defence"], s["std_attack"], s["std_defence"]
        self.standardised_attack, self.standardised_defence = s["standardised_attack"], s["standardised_defence"]
        self.attack_coefficients = s.get("attack_coefficients") if arr.covariates is not None else None
        self.defence_coefficients = s.get("defence_coefficients") if arr.covariates is not None else None
        if self.model == "neutral_wc":
            self.confederation_strength = s["confederation_strength_decentered"]  # LocScaleReparam of N(0,1): identity
            self.conferences, self._conferences_dict = self._meta["conferences"], self._meta["conferences_dict"]
        return self

    def _samples(self):
        d = {"attack": self.attack, "defence": self.defence, "corr_coef": self.corr_coef}
        for nm in self._effects:
            d[nm] = getattr(self, nm)
        return d

    def _fixture_extras(self, n, neutral_venue=0, **kw):
        nv = np.asarray(_str_to_list(neutral_venue)[0], dtype=np.uint8)
        return {"neutral_venue": np.resize(nv, n).astype(np.uint8)}

    def predict_score_proba(self, home_team, away_team, home_goals, away_goals, neutral_venue):
        return super().predict_score_proba(home_team, away_team, home_goals, away_goals, neutral_venue=neutral_venue)

    def predict_score_grid_proba(self, home_team, away_team, neutral_venue, max_goals: int = MAX_GOALS):
        return super().predict_score_grid_proba(home_team, away_team, max_goals=max_goals, neutral_venue=neutral_venue)

    def predict_outcome_proba(self, home_team, away_team, neutral_venue, knockout: bool = False, max_goals: int = MAX_GOALS):
        return super().predict_outcome_proba(home_team, away_team, max_goals=max_goals, knockout=knockout,
                                             neutral_venue=neutral_venue)

    def predict_score_n_proba(self, n, team, opponent, home: bool = True, neutral_venue: int = 0, max_goals: int = MAX_GOALS):
        return self._n_proba(n, team, opponent, home, max_goals, True, neutral_venue=neutral_venue)

    def sample_score(self, home_team, away_team, neutral_venue, num_samples: int = 1, random_state: Optional[int] = None,
                     max_goals: int = MAX_GOALS):
        return self._sample_score(home_team, away_team, num_samples, random_state, max_goals, neutral_venue=neutral_venue)

    def sample_outcome(self, home_team, away_team, neutral_venue, knockout: bool = False, num_samples: int = 1,
                       random_state: Optional[int] = None, max_goals: int = MAX_GOALS):
        return self._sample_outcome(home_team, away_team, num_samples, random_state, max_goals, knockout=knockout,
                                    neutral_venue=neutral_venue)

    def add_new_team(self, team_name: str, team_covariates: Optional[np.ndarray] = None) -> None:
        """``neutral_dixon_coles.py:490-560`` / ``neutral_dixon_coles_WC.py:476-546``: the pair from the correlated prior,
        the four venue effects from their own normal priors."""
        attack, defence = self._new_team_pair(team_name, team_covariates)
        cols = {"attack": attack, "defence": defence}
        for nm in self._effects:
            cols[nm] = np.random.normal(loc=getattr(self, "mean_" + nm), scale=getattr(self, "std_" + nm))
        self._append_team(team_name, cols)

    def predict_concede_n_proba(self, n, team, opponent, home: bool = True, neutral_venue: int = 0,
                                max_goals: int = MAX_GOALS):
        return self._n_proba(n, team, opponent, home, max_goals, False, neutral_venue=neutral_venue)


class NeutralDixonColesMatchPredictorWC(NeutralDixonColesMatchPredictor):
    """``bpl/neutral_dixon_coles_WC.py:31-968``: neutral model + per-confederation strength."""
    model = "neutral_wc"

    def fit(self, training_data, epsilon: float = 0.0, rescale_weights: bool = False, random_state: int = 42,
            num_warmup: int = 500, num_samples: int = 1000, mcmc_kwargs: Optional[Dict[str, Any]] = None,
            run_kwargs: Optional[Dict[str, Any]] = None):
        return super().fit(training_data, epsilon, rescale_weights, random_state, num_warmup, num_samples, mcmc_kwargs,
                           run_kwargs)

    def _samples(self):
        d = super()._samples()
        d["confederation_strength"] = self.confederation_strength
        return d

    def _fixture_extras(self, n, home_conf=None, away_conf=None, neutral_venue=0, **kw):
        out = super()._fixture_extras(n, neutral_venue=neutral_venue)
        for key, val in (("home_conf", home_conf), ("away_conf", away_conf)):
            v = _str_to_list(val)[0]
            if isinstance(v[0], str):
                v = [self._conferences_dict[c] for c in v]
            out[key] = np.resize(np.asarray(v, dtype=np.uint8), n)
        return out

    def predict_score_proba(self, home_team, away_team, home_conf, away_conf, home_goals, away_goals, neutral_venue):
        return _BplxPredictor.predict_score_proba(self, home_team, away_team, home_goals, away_goals, home_conf=home_conf,
                                                  away_conf=away_conf, neutral_venue=neutral_venue)

    def predict_score_grid_proba(self, home_team, away_team, home_conf, away_conf, neutral_venue, max_goals: int = MAX_GOALS):
        return _BplxPredictor.predict_score_grid_proba(self, home_team, away_team, max_goals=max_goals, home_conf=home_conf,
                                                       away_conf=away_conf, neutral_venue=neutral_venue)

    def predict_outcome_proba(self, home_team, away_team, home_conf, away_conf, neutral_venue, knockout: bool = False,
                              max_goals: int = MAX_GOALS):
        return _BplxPredictor.predict_outcome_proba(self, home_team, away_team, max_goals=max_goals, knockout=knockout,
                                                    home_conf=home_conf, away_conf=away_conf, neutral_venue=neutral_venue)

    def sample_score(self, home_team, away_team, home_conf, away_conf, neutral_venue, num_samples: int = 1,
                     random_state: Optional[int] = None, max_goals: int = MAX_GOALS):
        return self._sample_score(home_team, away_team, num_samples, random_state, max_goals, home_conf=home_conf,
                                  away_conf=away_conf, neutral_venue=neutral_venue)

    def sample_outcome(self, home_team, away_team, home_conf, away_conf, neutral_venue, knockout: bool = False,
                       num_samples: int = 1, random_state: Optional[int] = None, max_goals: int = MAX_GOALS):
        return self._sample_outcome(home_team, away_team, num_samples, random_state, max_goals, knockout=knockout,
                                    home_conf=home_conf, away_conf=away_conf, neutral_venue=neutral_venue)

    def _conf_kw(self, team_conf, opponent_conf, home, neutral_venue):
        hc, ac = (team_conf, opponent_conf) if home else (opponent_conf, team_conf)
        return dict(home_conf=hc, away_conf=ac, neutral_venue=neutral_venue)

    def predict_score_n_proba(self, n, team, opponent, team_conf, opponent_conf, home: bool = True, neutral_venue: int = 0,
                              max_goals: int = MAX_GOALS):
        return self._n_proba(n, team, opponent, home, max_goals, True, **self._conf_kw(team_conf, opponent_conf, home, neutral_venue))

    def predict_concede_n_proba(self, n, team, opponent, team_conf, opponent_conf, home: bool = True, neutral_venue: int = 0,
                                max_goals: int = MAX_GOALS):
        return self._n_proba(n, team, opponent, home, max_goals, False, **self._conf_kw(team_conf, opponent_conf, home, neutral_venue))


class DynamicNeutralDixonColesMatchPredictor(_BplxPredictor):
    """``bpl/dynamic_dixon_coles.py:23-334`` -- ``fit`` only.  The reference's predict methods index a Python list with a
    tuple and cannot run, so they are not mirrored.  ``walk="intended"`` (default) makes attack / defence the
    cumulative random walk the model describes; ``walk="as_written"`` reproduces ``:192-218`` literally.
    Gameweeks are 0-based and G = max + 1 (the reference passes ``max``)."""
    model = "dynamic"

    def __init__(self, walk: str = "intended"):
        super().__init__()
        if walk not in ("intended", "as_written"):
            raise ValueError(walk)
        self.walk = walk

    def fit(self, training_data, random_state: int = 42, num_warmup: int = 500, num_samples: int = 1000,
            mcmc_kwargs: Optional[Dict[str, Any]] = None, run_kwargs: Optional[Dict[str, Any]] = None):
        kw = dict(mcmc_kwargs or {})
        num_chains = int(kw.pop("num_chains", 1))
        thin = int(kw.pop("thin", kw.pop("thinning", 1)))
        kw.pop("chain_method", None)
        kw.pop("progress_bar", None)
        arr, meta = bdata.prepare("dynamic", training_data)
        arr.as_written = self.walk == "as_written"
        self.teams, self._teams_dict = list(meta["teams"]), meta["teams_dict"]
        self.problem = p = Problem(arr)
        G, T = arr.num_gameweeks, arr.num_teams
        flat_dev, cc = _run_chains(p, num_chains, random_state, num_warmup, num_samples, thin, kw, self)
        flat = flat_dev.cpu().numpy()
        fi = np.finfo(np.float32)
        s = {}
        for name, (off, cnt, tr) in p.layout.items():
            x = flat[:, off:off + cnt]
            if tr == "exp":
                x = np.exp(x)
            elif tr == "sigmoid":
                x = np.clip(1.0 / (1.0 + np.exp(-x)), fi.tiny, 1.0 - fi.eps)
            s[name] = x.astype(np.float32)
        S = flat.shape[0]
        gt = lambda k: s[k].reshape(S, G, T)  # noqa: E731
        self.corr_coef = cc.cpu().numpy()
        self.u, self.rho = gt("u"), 2.0 * gt("u") - 1.0
        self.standardised_attack, self.standardised_defence = gt("standardised_attack"), gt("standardised_defence")
        self.mean_defence = s["mean_defence"][:, 0]
        self.std_attack, self.std_defence = s["std_attack"], s["std_defence"]  # [S, G]
        for nm in ("home_attack", "away_attack", "home_defence", "away_defence"):
            setattr(self, "mean_" + nm, s["mean_" + nm])
            setattr(self, "std_" + nm, s["std_" + nm])
            setattr(self, nm, s["mean_" + nm][:, :, None] + s["std_" + nm][:, :, None] * gt(nm + "_decentered"))
        am, dm = np.zeros((S, 1), np.float32), self.mean_defence[:, None]
        if arr.covariates is not None:
            Xs = arr.covariates.astype(np.float32)
            am, dm = s["attack_coefficients"] @ Xs.T, dm + s["defence_coefficients"] @ Xs.T
            self.attack_coefficients, self.defence_coefficients = s["attack_coefficients"], s["defence_coefficients"]
        else:
            self.attack_coefficients = self.defence_coefficients = None
        step_a = self.standardised_attack * self.std_attack[:, :, None]
        step_d = self.standardised_defence * self.std_defence[:, :, None]
        if self.walk == "intended":
            att, dfn = am[:, None, :] + np.cumsum(step_a, axis=1), dm[:, None, :] + np.cumsum(step_d, axis=1)
        else:  # the recorded deterministics attack_j = 0 + z[j] std[j] for j >= 1
            att, dfn = step_a.copy(), step_d.copy()
            att[:, 0] += am
            dfn[:, 0] += dm
        self.attack = [att[:, j] for j in range(G)]  # the reference stores lists of [S, T] arrays (:308-309)
        self.defence = [dfn[:, j] for j in range(G)]
        return self
