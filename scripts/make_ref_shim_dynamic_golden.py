"""Golden vectors of the reference's DYNAMIC `_model` source (bpl/dynamic_dixon_coles.py:63-247), run in the build
container under oracle/ref_shim.py -- the a7 pin.

The reference's `fit()` (":250-296") is called for real and `MCMC.run` intercepted.  As written it passes
`num_gameweeks = max(gameweek)` (":287"), so the last gameweek indexes one row past its [G, T] arrays; JAX's gather
clamps such an index to the last row, torch (the stand-in) raises.  Two traces of the UNMODIFIED `_model` are stored:

  refdyn_as_fitted.npz   the arguments `fit()` handed over, with JAX's clamp made explicit (gameweek -> min(gameweek, G-1),
                         G = max(gameweek)): exactly what the reference computes on this data
  refdyn_g_plus_1.npz    the same model with num_gameweeks = max(gameweek) + 1 (every gameweek has its own row): the
                         reading the product implements

In both, attack / defence never reach the rates (the `.at[].set` results are discarded, ":192-218"): this is
the product's BPLX_FLAG_DYNAMIC_AS_WRITTEN mode.  Positions: radius 0.3 ... 1.5 in this repo's flat layout."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H  # noqa: F401  (before the reference is put on sys.path: it has a `tests` package too)
from oracle import ref_shim
ref_shim.install()
from bpl.dynamic_dixon_coles import DynamicNeutralDixonColesMatchPredictor as RefDynamic
from oracle import models as om

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def data(seed=11, T=6, weeks=5, M=160, K=0):
    rng = np.random.default_rng(seed)
    names = [chr(ord("A") + i) for i in range(T)]
    h = rng.integers(0, T, M)
    a = (h + rng.integers(1, T, M)) % T
    td = {"home_team": [names[i] for i in h], "away_team": [names[i] for i in a],
          "home_goals": rng.poisson(1.0, M), "away_goals": rng.poisson(0.9, M),
          "gameweek": rng.integers(0, weeks, M), "neutral_venue": (rng.random(M) < 0.3).astype(np.int64)}
    td["gameweek"][:weeks] = np.arange(weeks)  # every gameweek occurs
    if K:
        td["team_covariates"] = {n: rng.normal(size=K) * (1.0 + np.arange(K)) for n in names}
    return td, h, a


for tag, K in (("", 0), ("_cov", 2)):
    td, h, a = data(K=K)
    cap = ref_shim.capture_fit(RefDynamic(), td)
    names = ["home_team", "away_team", "gameweek", "num_teams", "num_gameweeks", "home_goals", "away_goals", "neutral_venue"]
    args = dict(zip(names, cap.args))
    T, G_fit = int(args["num_teams"]), int(args["num_gameweeks"])
    assert G_fit == int(np.max(td["gameweek"]))  # dynamic_dixon_coles.py:287
    for case, G, gw in (("as_fitted", G_fit, np.minimum(np.asarray(args["gameweek"]), G_fit - 1)),
                        ("g_plus_1", G_fit + 1, np.asarray(args["gameweek"]))):
        cap2 = ref_shim.Captured(cap.model, (args["home_team"], args["away_team"], gw, T, G, args["home_goals"],
                                             args["away_goals"], args["neutral_venue"]), cap.kwargs)
        layout = om.site_layout("dynamic", T, K, 0, G)
        offs = {k: v for k, v in om.layout_offsets(layout).items() if k != "__D__"}
        D = om.num_params("dynamic", T, K, 0, G)
        rng = np.random.default_rng(7)
        radii = np.array([0.3, 0.6, 1.0, 1.5])
        theta = rng.uniform(-1.0, 1.0, (len(radii), D)) * radii[:, None]
        lps, grads = [], []
        for i in range(len(radii)):
            vals = {}
            for name, (o, shape, _tr) in offs.items():
                cnt = int(np.prod(shape)) if shape else 1
                if cnt:
                    vals[name] = theta[i, o:o + cnt].reshape(shape)
            lp, g, det = ref_shim.log_density(cap2, vals)
            flat = np.zeros(D)
            for name, (o, shape, _tr) in offs.items():
                cnt = int(np.prod(shape)) if shape else 1
                if cnt:
                    flat[o:o + cnt] = np.asarray(g[name]).reshape(-1)
            lps.append(lp)
            grads.append(flat)
        X = None
        if K:  # the model standardises the covariates itself (":103-105"); the product's host prep does the same in float32
            Xr = np.array([td["team_covariates"][n] for n in sorted(td["team_covariates"])])
            X = ((Xr - Xr.mean(0)) / Xr.std(0)).astype(np.float32)
        extra = {} if X is None else {"covariates": X}
        np.savez_compressed(os.path.join(OUT, f"refdyn_{case}{tag}.npz"), theta=theta, lp=np.array(lps), grad=np.array(grads),
                            home_team=h.astype(np.uint16), away_team=a.astype(np.uint16),
                            home_goals=np.asarray(td["home_goals"]).astype(np.uint8),
                            away_goals=np.asarray(td["away_goals"]).astype(np.uint8),
                            neutral_venue=np.asarray(td["neutral_venue"]).astype(np.uint8), gameweek=gw.astype(np.int32),
                            num_teams=T, num_gameweeks=G, **extra)
        # the oracle (as-written walk) on the same positions, right away
        from bpl_next_b200 import data as bdata
        arr = bdata.MatchArrays(model="dynamic", num_teams=T, home_team=h.astype(np.uint16), away_team=a.astype(np.uint16),
                                home_goals=np.asarray(td["home_goals"]).astype(np.uint8),
                                away_goals=np.asarray(td["away_goals"]).astype(np.uint8),
                                neutral_venue=np.asarray(td["neutral_venue"]).astype(np.uint8), gameweek=gw.astype(np.int32),
                                num_gameweeks=G, covariates=X, as_written=True)
        lo, go, _ = om.log_density_and_grad(H.to_oracle(arr), theta)
        e_lp = np.max(np.abs(lo - np.array(lps)) / np.abs(np.array(lps)))
        e_g = np.max(np.abs(go - np.array(grads)) / np.abs(np.array(grads)).max(axis=1, keepdims=True))
        print(f"refdyn_{case}{tag}: D={D} G={G}  lp ref {lps[0]:.6f} oracle {lo[0]:.6f}  max rel err lp {e_lp:.2e} grad {e_g:.2e}")
