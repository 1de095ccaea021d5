#!/usr/bin/env python
"""bench.py -- log-density+gradient match-evaluations per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A *step* is one pass of the hot path over one batch: one `bplx_logdensity_fwdbwd` call for all the
chains of this GPU.  N=1 workload = BASELINE.json configs[1] (ExtendedDixonColes, 5 synthetic
seasons, 4,096 vectorised chains); N>1 = the same per-GPU workload on every rank (weak scaling,
chains are independent: no data-path collective).  `value` times the kernel with inputs resident
in HBM (CUDA graph of K launches over rotating buffer sets larger than L2, CUDA events on the
launch stream); `e2e` goes through the host-buffer C-ABI call with pinned host arrays, copies
inside the timed region.  `--impl reference` times the CPU restatement of the reference
(oracle/models.py, float32, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_BYTES = 126e6
FLOPS_PER_EVAL = {"dixon_coles": 26, "extended": 30, "neutral": 40, "neutral_wc": 46, "dynamic": 40}
METRIC = "logdensity_grad_match_evals_per_s"
UNIT = "match-evals/s"


# ------------------------------------------------------------------------------------------------
# workloads (synthetic; oracle/datasets.py only generates inputs here)
# ------------------------------------------------------------------------------------------------
def workload(name):
    from bpl_next_b200 import data as bdata
    from oracle import datasets

    if name == "cfg2":
        arr, _ = bdata.prepare("extended", datasets.config_2(), epsilon=0.01)
        desc = "configs[1]: ExtendedDixonColes T=20, 5 seasons M=1900, K=3 covariates, eps=0.01"
        return arr, 4096, desc
    if name == "cfg3":
        arr, _ = bdata.prepare("neutral_wc", datasets.config_3(), epsilon=0.1)
        desc = "configs[2]: NeutralDixonColesWC T=220, M=40000 weighted, Cf=6"
        return arr, 32768, desc
    if name == "cfg4":
        arr, _ = bdata.prepare("dynamic", datasets.config_4())
        desc = "configs[3]: DynamicDixonColes T=20, 30 seasons M=11400, G=30 (intended random walk)"
        return arr, 8192, desc
    if name == "cfg1":
        arr, _ = bdata.prepare("dixon_coles", datasets.dummy_data())
        return arr, 4096, "configs[0] data: DixonColes T=20 M=380"
    raise ValueError(name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written) + FFMA / smem peaks measured now."""
    out = {"hbm_gbs": 6650.0, "hbm_source": "fallback"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            out["hbm_gbs"] = float(json.load(f)["hbm_gbs"])
        out["hbm_source"] = "measured (MEASURED_PEAKS.json)"
    lib = ctypes.CDLL(os.path.join(ROOT, "bpl_next_b200", "lib", "libbplx_bench.so"))
    lib.bplxbench_fp32_peak.restype = ctypes.c_double
    lib.bplxbench_smem_peak.restype = ctypes.c_double
    out["fp32_tflops"] = float(lib.bplxbench_fp32_peak(2000, 5))
    out["smem_tbs"] = float(lib.bplxbench_smem_peak(2000, 5))
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def time_logdensity(problem, C, steps, warmup, radius, seed, use_graph=True, target_s=1.0):
    """Median device time of a K-step region (ms) over several repetitions."""
    import torch

    D = problem.D
    set_bytes = 2 * C * D * 4
    nb = max(2, int(math.ceil(1.5 * L2_BYTES / set_bytes)))
    nb = min(nb, max(2, steps))
    g = torch.Generator(device="cuda").manual_seed(seed)
    thetas = [(torch.rand((D, C), generator=g, device="cuda", dtype=torch.float32) * 2 - 1) * radius for _ in range(nb)]
    grads = [torch.empty((D, C), device="cuda", dtype=torch.float32) for _ in range(nb)]
    lps = [torch.empty(C, device="cuda", dtype=torch.float32) for _ in range(nb)]
    ccs = [torch.empty(C, device="cuda", dtype=torch.float32) for _ in range(nb)]
    problem.workspace(C)
    stream = torch.cuda.Stream()

    def launch(k):
        i = k % nb
        problem.logdensity(thetas[i], chain_minor=True, lp=lps[i], grad=grads[i], corr_coef=ccs[i], stream=stream)

    with torch.cuda.stream(stream):
        for k in range(warmup):
            launch(k)
        stream.synchronize()
        graph = None
        if use_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                for k in range(steps):
                    launch(k)

        def region():
            if graph is not None:
                graph.replay()
            else:
                for k in range(steps):
                    launch(k)

        region()  # warm the graph / caches once
        stream.synchronize()
        times = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start = time.perf_counter()
        while True:
            e0.record(stream)
            region()
            e1.record(stream)
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
            if time.perf_counter() - t_start > target_s or len(times) >= 400:
                break
    finite = bool(torch.isfinite(lps[0]).all().item())
    return float(np.median(times)), len(times), nb, set_bytes, finite


def time_e2e(problem, C, steps, warmup, radius, seed):
    """Host-buffer entry point: pinned numpy in, pinned numpy out, copies inside the timed region."""
    import torch

    D = problem.D
    rng = np.random.default_rng(seed)
    nb = 4
    th = [torch.empty((C, D), dtype=torch.float32).pin_memory() for _ in range(nb)]
    for t in th:
        t.numpy()[:] = rng.uniform(-radius, radius, (C, D)).astype(np.float32)
    lp = torch.empty(C, dtype=torch.float32).pin_memory()
    gr = torch.empty((C, D), dtype=torch.float32).pin_memory()
    cc = torch.empty(C, dtype=torch.float32).pin_memory()
    for k in range(max(warmup, 3)):
        problem.logdensity_host(th[k % nb].numpy(), lp=lp.numpy(), grad=gr.numpy(), corr_coef=cc.numpy())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        problem.logdensity_host(th[k % nb].numpy(), lp=lp.numpy(), grad=gr.numpy(), corr_coef=cc.numpy())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return dt, C * D * 4, C * D * 4 + 2 * C * 4


def cpu_port(arr, C_sample, min_seconds, radius, seed, max_calls=1000):
    """The oracle restatement (float32, autograd) on all host threads; returns evals/s and details."""
    import torch

    from oracle import models as om
    from tests import helpers as H

    d = H.to_oracle(arr)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    D = om.num_params(arr.model, arr.num_teams, arr.num_covariates, arr.num_conferences, arr.num_gameweeks)
    theta = np.random.default_rng(seed).uniform(-radius, radius, (C_sample, D)).astype(np.float32)
    om.log_density_and_grad(d, theta, dtype=torch.float32)  # warm-up
    calls, t0 = 0, time.perf_counter()
    per_call = []
    while True:
        t1 = time.perf_counter()
        om.log_density_and_grad(d, theta, dtype=torch.float32)
        per_call.append(time.perf_counter() - t1)
        calls += 1
        if time.perf_counter() - t0 >= min_seconds or calls >= max_calls:
            break
    dt = time.perf_counter() - t0
    return {"value": C_sample * arr.num_matches * calls / dt, "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "port", "per_call_s": float(np.median(per_call)), "calls": calls,
            "sample": f"{C_sample} chains x {arr.num_matches} matches per call, {calls} calls, "
                      f"oracle/models.py float32 value+autograd gradient (restatement of the reference, not its binary)"}


def run_reference(args, rank, world):
    """Reference arm: the CPU restatement timed on the host cores (rank 0 only)."""
    if rank != 0:
        return
    arr, C, desc = workload(args.workload)
    C_sample = args.cpu_chains
    import torch
    from oracle import models as om
    from tests import helpers as H

    d = H.to_oracle(arr)
    torch.set_num_threads(os.cpu_count() or 1)
    D = om.num_params(arr.model, arr.num_teams, arr.num_covariates, arr.num_conferences, arr.num_gameweeks)
    theta = np.random.default_rng(args.seed + 1).uniform(-args.radius, args.radius, (C_sample, D)).astype(np.float32)
    for _ in range(max(args.warmup, 1)):
        om.log_density_and_grad(d, theta, dtype=torch.float32)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        om.log_density_and_grad(d, theta, dtype=torch.float32)
    dt = time.perf_counter() - t0
    value = C_sample * arr.num_matches * args.steps / dt
    sample = (f"each step = {C_sample} chains x {arr.num_matches} matches (bounded sample of {C} chains/GPU), "
              "oracle/models.py float32 value+autograd gradient on all host threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc + f", {C} chains/GPU", "chains_per_gpu": C, "matches": arr.num_matches,
                       "teams": arr.num_teams, "params": D, "theta": f"U(-{args.radius},{args.radius})"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def extras(args, pk):
    """Secondary workloads on one GPU (reported inside the main JSON line, not bench lines of their own)."""
    import torch

    from bpl_next_b200 import Problem, score_grid
    from oracle import datasets

    out = []
    # configs[1] again with theta near the mode, U(-0.2, 0.2): what a sampler past its first warm-up steps sends.  No rate
    # of any chain is near the clip at 15, so K1 takes the clip-free forms of the same arithmetic.
    arr, C, desc = workload("cfg2")
    p = Problem(arr)
    ms, reps, nb, set_bytes, finite = time_logdensity(p, C, 50, 5, 0.2, args.seed + 9, use_graph=True, target_s=0.5)
    evals = C * arr.num_matches * 50 / (ms * 1e-3)
    out.append({"workload": desc + f", {C} chains on 1 GPU, theta U(-0.2,0.2) (typical-set inputs: clip-free forms)",
                "metric": METRIC, "value": evals, "unit": UNIT, "ms_per_call": ms / 50, "finite": finite,
                "fp32_frac": FLOPS_PER_EVAL[arr.model] * evals / 1e12 / pk["fp32_tflops"]})
    p.close()
    del p
    # configs[2]: NeutralWC, 32,768 chains on one GPU (inputs 2 x 175 MB > L2)
    arr, C, desc = workload("cfg3")
    p = Problem(arr)
    ms, reps, nb, set_bytes, finite = time_logdensity(p, C, 3, 2, args.radius, args.seed + 3, use_graph=False,
                                                      target_s=0.5)
    evals = C * arr.num_matches * 3 / (ms * 1e-3)
    out.append({"workload": desc + f", {C} chains on 1 GPU", "metric": METRIC, "value": evals, "unit": UNIT,
                "ms_per_call": ms / 3, "finite": finite, "plan": p.stats(),
                "fp32_frac": FLOPS_PER_EVAL["neutral_wc"] * evals / 1e12 / pk["fp32_tflops"],
                # one float2 table row per chain and list entry: the algorithmic shared-memory bytes against the LDS.64
                # peak measured in this run (ncu counts 1.3x more wavefronts than that: the broadcast entry reads, bank
                # conflicts -- profiles/r01_k1_final_cfg3.md)
                "smem_frac": 8.0 * C * (p.stats()["entries1_padded"] + p.stats()["entries2_padded"]) / (ms / 3 * 1e-3) / 1e12
                             / pk["smem_tbs"]})
    # the few-chain ("streaming") regime on the same data: one CTA walks the whole static plan
    st = p.stats()
    plan_bytes = 8 * (st["entries1_padded"] + st["entries2_padded"]) + 16 * 2600
    for Cs in (1, 32):
        ms1, _, _, _, fin1 = time_logdensity(p, Cs, 20, 3, args.radius, args.seed + 5, use_graph=True, target_s=0.3)
        out.append({"workload": desc + f", {Cs} chain(s): few-chain regime, one CTA", "metric": METRIC,
                    "value": Cs * arr.num_matches * 20 / (ms1 * 1e-3), "unit": UNIT, "ms_per_call": ms1 / 20, "finite": fin1,
                    "plan_stream_gbs": plan_bytes / (ms1 / 20 * 1e-3) / 1e9,
                    "note": "latency-bound by design: the plan (0.8 MB) is read once per CTA through the TMA ring; a "
                            "match-parallel streaming kernel (K1s) only pays above ~1e6 matches and is not built"})
    p.close()
    del p
    torch.cuda.empty_cache()
    # configs[3]: Dynamic, 8,192 chains (theta radius 0.5: a 30-step walk of U(-2,2) steps overflows float32 rates)
    arr, C, desc = workload("cfg4")
    p = Problem(arr)
    ms, reps, nb, set_bytes, finite = time_logdensity(p, C, 3, 2, 0.5, args.seed + 4, use_graph=False, target_s=0.5)
    evals = C * arr.num_matches * 3 / (ms * 1e-3)
    out.append({"workload": desc + f", {C} chains on 1 GPU, theta U(-0.5,0.5)", "metric": METRIC, "value": evals,
                "unit": UNIT, "ms_per_call": ms / 3, "finite": finite, "plan": p.stats(),
                "fp32_frac": FLOPS_PER_EVAL["dynamic"] * evals / 1e12 / pk["fp32_tflops"],
                "hbm_frac": (8.0 * p.D * C) / (ms / 3 * 1e-3) / 1e9 / pk["hbm_gbs"]})
    p.close()
    del p
    torch.cuda.empty_cache()
    # configs[0]: DixonColesMatchPredictor.fit on the 20-team / 380-match season -> ESS/s (whole fit, warm-up included)
    try:
        from bpl_next_b200 import DixonColesMatchPredictor, diagnostics as dg
        for chains, nw, ns in ((1, 500, 1000), (1024, 500, 250)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m = DixonColesMatchPredictor().fit(datasets.dummy_data(), num_warmup=nw, num_samples=ns,
                                               mcmc_kwargs={"num_chains": chains})
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            run = m.nuts_run
            ess = dg.effective_sample_size(run.samples)
            rhat = dg.split_rhat(run.samples) if ns >= 4 else None
            out.append({"workload": f"configs[0]: DixonColesMatchPredictor.fit T=20 M=380, {chains} chain(s), {nw} warmup + {ns} draws",
                        "metric": "ess_per_s", "value": float(ess.min().item()) / wall, "unit": "min-over-parameters bulk ESS / s",
                        "fit_wall_s": wall, "ess_min": float(ess.min().item()), "ess_median": float(ess.median().item()),
                        "rhat_max": None if rhat is None else float(rhat.max().item()),
                        "logdensity_launches": int(run.launches), "leapfrogs_total": int(run.num_leapfrog.sum()),
                        "divergences": int(run.num_divergent.sum()),
                        "match_evals_per_s_inside_fit": float(run.num_leapfrog.sum()) * 380 / wall})
            del m
    except Exception as e:  # never lose the main line
        out.append({"workload": "configs[0] fit", "error": repr(e)})
    # configs[4]: predictive grid S=16,384 x F=10,000 x 11x11 (whole job on one GPU here)
    s, fx = datasets.config_5()
    ds = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
    dfx = {k: torch.from_numpy(v).cuda() for k, v in fx.items()}
    S, F = s["attack"].shape[0], len(fx["home_team"])
    ws = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    grid = torch.empty((F, 11, 11), dtype=torch.float32, device="cuda")
    outc = torch.empty((F, 3), dtype=torch.float32, device="cuda")
    for _ in range(3):
        score_grid("neutral_wc", ds, dfx, 10, workspace=ws, grid=grid, outcome=outc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        e0.record()
        score_grid("neutral_wc", ds, dfx, 10, workspace=ws, grid=grid, outcome=outc)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    flops = S * F * (2 * 121 + 2 * 2 * 11 + 12)
    out.append({"workload": f"configs[4]: predict grid S={S} x F={F} x 11x11 on 1 GPU", "metric": "score_grid_ms",
                "value": ms, "unit": "ms", "sample_fixture_pairs_per_s": S * F / (ms * 1e-3),
                "fp32_frac": flops / (ms * 1e-3) / 1e12 / pk["fp32_tflops"],
                "outcome_sum_err": float((outc.sum(dim=1) - 1).abs().max().item())})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="bplx", choices=["bplx", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4"])
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the workload's)")
    ap.add_argument("--radius", type=float, default=2.0, help="theta ~ U(-radius, radius) (numpyro init_to_uniform)")
    ap.add_argument("--seed", type=int, default=1002)
    ap.add_argument("--cpu-chains", type=int, default=256, help="chains per call of the CPU restatement")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from bpl_next_b200 import Problem, _abi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (bpl_next_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    arr, C, desc = workload(args.workload)
    if args.chains:
        C = args.chains
    problem = Problem(arr)
    lib = _abi.lib()
    pk = peaks() if rank == 0 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -----------------------------------------------------------------
    barrier()
    launches0 = lib.bplx_launch_count()
    with ClockSampler(local) as clk:
        ms, reps, nb, set_bytes, finite = time_logdensity(problem, C, args.steps, args.warmup, args.radius,
                                                          args.seed + 1 + rank, use_graph=not args.no_graph)
    barrier()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches_per_region = args.steps  # one kernel per step (graph replays re-issue the captured launches)
    value = world * C * arr.num_matches * args.steps / (ms * 1e-3)

    # ---- end to end (host buffers) ---------------------------------------------------------------------
    barrier()
    dt, h2d, d2h = time_e2e(problem, C, args.steps, args.warmup, args.radius, args.seed + 7 + rank)
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * C * arr.num_matches * args.steps / float(t.item())
    barrier()

    if rank == 0:
        st = problem.stats()
        t_launch = ms * 1e-3 / args.steps
        model = arr.model
        flops = FLOPS_PER_EVAL[model] * C * arr.num_matches
        ent_bytes = (16 if model == "extended" else 8) * st["entries1_padded"] + 8 * st["entries2_padded"]
        hbm_bytes = ent_bytes + 8 * problem.D * C + 8 * C  # static plan + theta in + grad out + lp, corr_coef
        smem_bytes = 8 * C * (st["entries1_padded"] + st["entries2_padded"])  # one float2 row read per chain-entry
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = (json.load(f).get(args.workload) or {}).get("traffic")  # dram read + write bytes of one launch (ncu --set full)
        roofline = {
            "bound": "fp32", "kernel": "bplx::logdensity_kernel",
            "achieved": flops / t_launch / 1e12, "peak": pk["fp32_tflops"], "unit": "TFLOP/s",
            "frac": flops / t_launch / 1e12 / pk["fp32_tflops"],
            "peak_source": "FFMA issue peak measured in this run (bench_kernels/fp32_peak.cu); "
                           "algorithmic flops/eval from SURVEY.md 8(d)",
            "flops_per_eval": FLOPS_PER_EVAL[model], "traffic": traffic,
            "hbm": {"achieved": hbm_bytes / t_launch / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": hbm_bytes / t_launch / 1e9 / pk["hbm_gbs"], "algorithmic_bytes": hbm_bytes,
                    "peak_source": pk["hbm_source"]},
            "smem": {"achieved": smem_bytes / t_launch / 1e12, "peak": pk["smem_tbs"], "unit": "TB/s",
                     "frac": smem_bytes / t_launch / 1e12 / pk["smem_tbs"],
                     "peak_source": "LDS.64 peak measured in this run"},
        }
        # the CPU baseline is a reported number of the N=1 run only
        cpu = cpu_port(arr, args.cpu_chains, args.cpu_seconds, args.radius, args.seed + 1) if world == 1 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc + f", {C} chains/GPU", "chains_per_gpu": C, "matches": arr.num_matches,
                       "teams": arr.num_teams, "params": problem.D, "theta": f"U(-{args.radius},{args.radius})",
                       "layout": "chain-minor [D, C] resident in HBM",
                       "l2": f"rotating {nb} buffer sets of {set_bytes / 1e6:.1f} MB (> 126 MB L2)",
                       "timing": f"{'CUDA graph of' if not args.no_graph else ''} {args.steps} launches, CUDA events on "
                                 f"the launch stream, median of {reps} repetitions, max over ranks",
                       "plan": st},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "bplx_logdensity_fwdbwd_host (pinned numpy in/out, chain-major [C, D])"},
            "gpu_launches": launches_per_region,
            "launch_counter_total": int(lib.bplx_launch_count() - launches0),
            "finite": finite,
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        if world == 1 and not args.no_extras:
            try:
                line["extra_workloads"] = extras(args, pk)
            except Exception as e:  # extras must never lose the main line
                line["extra_workloads_error"] = repr(e)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
