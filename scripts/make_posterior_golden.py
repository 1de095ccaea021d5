"""Posterior summary of DixonColes on the reference's dummy season from an INDEPENDENT sampler: plain HMC (fixed number
of leapfrogs, Metropolis correction, diagonal mass from warm-up) on the CPU oracle density in float64 -- no code shared
with the CUDA kernels or the GPU NUTS.  Writes tests/golden/posterior_dixon_coles.npz (means, sds, Monte-Carlo standard
errors of attack / defence / home_advantage / the scales / corr_coef); tests/test_nuts_gpu.py compares a GPU fit with it.
Takes a few minutes."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import datasets, models as om
from tests import helpers as H
from bpl_next_b200 import diagnostics as dg

torch.set_num_threads(os.cpu_count() or 8)
MODEL = sys.argv[1] if len(sys.argv) > 1 else "dixon_coles"  # dixon_coles | extended | neutral_wc | tw
rng = np.random.default_rng(123)
C, L = 64, 12


def run_hmc(d, D, N=900):
    """Plain HMC (L leapfrogs, Metropolis correction, diagonal mass from warm-up) on the float64 oracle density."""
    def U(theta):
        return om.log_density_and_grad(d, theta)

    def hmc_step(theta, lp, g, eps, inv_mass):
        r0 = rng.normal(size=theta.shape) / np.sqrt(inv_mass)
        th, r, gg = theta.copy(), r0.copy(), g.copy()
        e = eps[:, None] * rng.uniform(0.8, 1.2, (len(eps), 1))
        lp1 = lp
        for _ in range(L):
            r = r + 0.5 * e * gg
            th = th + e * inv_mass * r
            lp1, gg, _cc = U(th)
            r = r + 0.5 * e * gg
        h0 = -lp + 0.5 * (inv_mass * r0 * r0).sum(1)
        h1 = -lp1 + 0.5 * (inv_mass * r * r).sum(1)
        dh = h0 - h1
        dh = np.where(np.isfinite(dh), dh, -np.inf)
        acc_p = np.minimum(1.0, np.exp(np.minimum(dh, 0.0)))
        take = rng.uniform(size=len(eps)) < acc_p
        theta = np.where(take[:, None], th, theta)
        g = np.where(take[:, None], gg, g)
        lp = np.where(take, lp1, lp)
        return theta, lp, g, acc_p

    theta = rng.uniform(-0.5, 0.5, (C, D))
    lp, g, _ = U(theta)
    eps = np.full(C, 0.05)
    inv_mass = np.ones(D)
    t0 = time.time()
    hist = []
    for stage, iters in (("step", 200), ("mass", 200), ("step2", 200)):
        for it in range(iters):
            theta, lp, g, acc = hmc_step(theta, lp, g, eps, inv_mass)
            eps = np.clip(eps * np.exp(0.08 * (acc - 0.8) * (1.0 if it < iters * 0.8 else 0.3)), 1e-4, 1.0)
            if stage == "mass" and it >= 50:
                hist.append(theta.copy())
        if stage == "mass":
            inv_mass = np.concatenate(hist, 0).var(axis=0) + 1e-3
            eps = np.full(C, np.median(eps) * 3.0)
        print(stage, "done, median eps %.4f mean accept %.2f, %.0f s" % (np.median(eps), acc.mean(), time.time() - t0), flush=True)
    draws = np.zeros((N, C, D))
    for it in range(N):
        theta, lp, g, acc = hmc_step(theta, lp, g, eps, inv_mass)
        draws[it] = theta
    print("sampling done, mean accept %.2f, %.0f s" % (acc.mean(), time.time() - t0), flush=True)
    return draws


if MODEL == "tw":
    # tests/test_extended_dixon_coles.py:28-47 of the reference asserts that doubling epsilon scales the posterior-mean
    # attack gap of `timed_dummy_data` by > 1.5 -- on ONE chain of 1000 draws.  The converged value under this density:
    out = {}
    for eps_w in (1.0, 2.0):
        arr = H.from_training_data("extended", datasets.timed_dummy_data(), epsilon=eps_w)
        d = H.to_oracle(arr)
        D = om.num_params("extended", arr.num_teams, 0, 0)
        offs = om.layout_offsets(om.site_layout("extended", arr.num_teams, 0, 0))
        draws = run_hmc(d, D, N=1500)
        o_sa, o_za = offs["std_attack"][0], offs["standardised_attack"][0]
        attack = np.exp(draws[:, :, o_sa:o_sa + 1]) * draws[:, :, o_za:o_za + 2]  # no covariates: prior mean 0
        gap = attack[:, :, 1] - attack[:, :, 0]  # [N, C]
        x = torch.from_numpy(np.ascontiguousarray(gap[:, None, :])).float()
        ess = float(dg.effective_sample_size(x).numpy()[0])
        out[f"gap_eps{int(eps_w)}_mean"] = gap.mean()
        out[f"gap_eps{int(eps_w)}_mcse"] = gap.std() / np.sqrt(ess)
        out[f"gap_eps{int(eps_w)}_rhat"] = float(dg.split_rhat(x).numpy()[0])
        print(f"eps {eps_w}: attack gap {gap.mean():.4f} +- {gap.std() / np.sqrt(ess):.4f} (ess {ess:.0f}, rhat {out[f'gap_eps{int(eps_w)}_rhat']:.3f})")
    r = out["gap_eps2_mean"] / out["gap_eps1_mean"]
    out["ratio"] = r
    out["ratio_mcse"] = abs(r) * np.sqrt((out["gap_eps1_mcse"] / out["gap_eps1_mean"]) ** 2 + (out["gap_eps2_mcse"] / out["gap_eps2_mean"]) ** 2)
    print(f"ratio {r:.4f} +- {out['ratio_mcse']:.4f}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "posterior_time_weighting.npz"), **out)
    sys.exit(0)

if MODEL in ("dixon_coles", "extended"):
    arr = H.from_training_data(MODEL, datasets.dummy_data())
else:
    arr = H.from_training_data("neutral_wc", datasets.neutral_dummy_data(), epsilon=0.2)
d = H.to_oracle(arr)
T = arr.num_teams
Cf = arr.num_conferences or 0
D = om.num_params(MODEL, T, 0, Cf)
offs = om.layout_offsets(om.site_layout(MODEL, T, 0, Cf))
N = 900
draws = run_hmc(d, D, N)


def site(name):
    o, shape, _ = offs[name]
    n = int(np.prod(shape)) if shape else 1
    return draws[:, :, o:o + n]


std_a, std_d = np.exp(site("std_attack")), np.exp(site("std_defence"))
if MODEL == "dixon_coles":
    q = {"attack": std_a * site("attack_decentered"), "defence": site("mean_defence") + std_d * site("defence_decentered"),
         "home_advantage": site("home_advantage"), "std_attack": std_a, "std_defence": std_d}
elif MODEL == "extended":
    sig = 1.0 / (1.0 + np.exp(-site("u")))
    q = {"attack": std_a * site("standardised_attack"), "defence": site("mean_defence") + std_d * site("standardised_defence"),
         "home_advantage": site("mean_home_advantage") + np.exp(site("std_home_advantage")) * site("home_advantage_decentered"),
         "rho": 2.0 * sig - 1.0, "std_attack": std_a, "std_defence": std_d}
else:
    q = {"attack": std_a * site("standardised_attack"), "defence": site("mean_defence") + std_d * site("standardised_defence"),
         "std_attack": std_a, "std_defence": std_d, "confederation_strength": site("confederation_strength_decentered")}
    for nm in ("home_attack", "away_attack", "home_defence", "away_defence"):
        q[nm] = site("mean_" + nm) + np.exp(site("std_" + nm)) * site(nm + "_decentered")
cc = np.stack([om.log_density_and_grad(d, draws[i])[2] for i in range(0, N, 3)], 0)  # every third draw is plenty
q["corr_coef"] = cc[:, :, None]
out = {}
for k, v in q.items():
    x = torch.from_numpy(np.ascontiguousarray(v.transpose(0, 2, 1)))  # [N, dims, C]
    ess = dg.effective_sample_size(x.float()).numpy()
    rhat = dg.split_rhat(x.float()).numpy()
    mean, sd = v.mean(axis=(0, 1)), v.reshape(-1, v.shape[-1]).std(axis=0)
    out[k + "_mean"], out[k + "_sd"], out[k + "_mcse"], out[k + "_rhat"] = mean, sd, sd / np.sqrt(ess), rhat
    print(f"{k:16s} rhat max {rhat.max():.3f} ess min {ess.min():.0f} mean[0] {mean[0]:+.4f} sd[0] {sd[0]:.4f} mcse[0] {(sd / np.sqrt(ess))[0]:.4f}")
np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"posterior_{MODEL}.npz"), **out)
