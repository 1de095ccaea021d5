"""torchrun, N ranks (one per GPU, NCCL): the two sharded paths against their single-GPU results.
  N1  predictive grid: samples sharded over ranks, all_reduce(sum) of [F, g, g]  == the grid from all samples on one GPU
  N2  fit: chains sharded over ranks (no collective in the hot loop); split-R-hat moments all-reduced
Writes gpurun_out/multi_gpu_check.json on rank 0."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpl_next_b200 import Problem, score_grid, parallel, nuts as bn, diagnostics as dg
from oracle import datasets
from tests import helpers as H

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = {"world": world}
# ---- N1 --------------------------------------------------------------------------------------------------------
S, F = 4096, 4100
s, fx = datasets.config_5(S=S, F=F)
start, cnt = parallel.shard(S, rank, world)
local_s = {k: torch.from_numpy(v[start:start + cnt]).cuda() for k, v in s.items()}
dfx = {k: torch.from_numpy(v).cuda() for k, v in fx.items()}
for _ in range(2):
    grid, outc = parallel.score_grid_sharded("neutral_wc", local_s, dfx, 10, S)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); grid, outc = parallel.score_grid_sharded("neutral_wc", local_s, dfx, 10, S); e1.record(); e1.synchronize()
t = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
full_s = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
ref, ref_out = score_grid("neutral_wc", full_s, dfx, 10)
torch.cuda.synchronize()
out["grid_max_abs_diff"] = float((grid - ref).abs().max().item())
out["outcome_max_abs_diff"] = float((outc - ref_out).abs().max().item())
out["grid_sharded_ms_max_over_ranks"] = float(t.item())
# ---- N1b: the overlapped, ranged variant; exchange over peer memory (bplx_peer_sum) against NCCL all_reduce -------------
for name, peer, graph in (("peer_graph", True, True), ("peer", True, False), ("nccl", False, False)):
    sg = parallel.ShardedScoreGrid("neutral_wc", local_s, dfx, 10, S, peer=peer, use_graph=graph)
    for _ in range(3):
        sg.run()
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(10):
        dist.barrier()
        ts.append(sg.run(timed=True))
    g2, o2 = sg.run()
    torch.cuda.synchronize()
    med = np.median(np.array(ts), axis=0)
    tt = torch.tensor(med, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    out[f"ranged_{name}"] = {"exchange": sg.exchange_kind, "uses_peer_memory": sg.peer is not None, "replayed_as_graph": sg.graphs is not None,
                             "ranges": len(sg.ranges),
                             "grid_max_abs_diff": float((g2 - ref).abs().max().item()),
                             "outcome_max_abs_diff": float((o2 - ref_out).abs().max().item()),
                             "total_ms": float(tt[0].item()), "compute_ms": float(tt[1].item()), "exposed_exchange_ms": float(tt[2].item())}
    chk = torch.stack([g2.double().sum(), o2.double().sum()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out[f"ranged_{name}"]["ranks_hold_identical_sums"] = bool(torch.equal(lo, hi))
# ---- N2 --------------------------------------------------------------------------------------------------------
arr = H.from_training_data("dixon_coles", datasets.dummy_data())
p = Problem(arr)
C_total = 256
c0, cn = parallel.shard(C_total, rank, world)
def potential(theta, lp, grad): p.logdensity(theta, chain_minor=True, lp=lp, grad=grad)
g = torch.Generator(device="cuda").manual_seed(100 + rank)
theta0 = torch.rand((p.D, cn), generator=g, device="cuda") * 4 - 2
t0 = time.perf_counter()
run = bn.sample(potential, theta0, num_warmup=300, num_samples=100, seed=5, chain_offset=c0)
torch.cuda.synchronize(); wall = time.perf_counter() - t0
x = run.samples.double()  # [N, D, C_local]
N = x.shape[0]
sx, sx2, sm2, n = parallel.allreduce_chain_moments(x.sum(dim=(0, 2)), (x * x).sum(dim=(0, 2)), (x.mean(dim=0) ** 2).sum(dim=1), cn)
mean = sx / (n * N)
var_total = sx2 / (n * N) - mean ** 2
between = sm2 / n - mean ** 2              # variance of the chain means
out.update({"fit_chains_total": n, "fit_wall_s_rank": wall, "home_advantage_mean": float(mean[0].item()),
            "max_between_over_total_var": float((between / var_total).max().item()),
            "local_rhat_max": float(dg.split_rhat(run.samples).max().item())})
# ---- fit through the predictor class: chains partitioned over the ranks, draws all-gathered at the end -------------
from bpl_next_b200 import DixonColesMatchPredictor
m = DixonColesMatchPredictor().fit(datasets.dummy_data(), num_warmup=200, num_samples=50, mcmc_kwargs={"num_chains": 64})
out["predictor_fit_draws"] = int(m.attack.shape[0])
chk = torch.tensor([float(m.attack.sum()), float(m.corr_coef.sum())], device="cuda", dtype=torch.float64)
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
out["predictor_fit_ranks_agree"] = bool(torch.equal(lo, hi))
probs = m.predict_outcome_proba("0", "1")
out["predictor_outcome_sum"] = float(probs["home_win"][0] + probs["draw"][0] + probs["away_win"][0])
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/multi_gpu_check.json", "w"))
    print(json.dumps(out))
dist.barrier(); dist.destroy_process_group()
