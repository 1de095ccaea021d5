// nuts.cu -- batched NUTS transition bookkeeping on the GPU, one thread per chain.
//
// The caller of the log-density kernel in the reference is numpyro's NUTS (`NUTS(self._model)` + `MCMC.run`,
// bpl/dixon_coles.py:100-116 and siblings): iterative tree doubling with multinomial sampling, biased progressive
// sampling at the top level, the generalised U-turn criterion on momentum sums, divergence at dH > 1000, dual
// averaging of the step size (t0 = 10, kappa = 0.75, gamma = 0.05, mu = log(10 eps)) and Stan-style windowed
// diagonal mass-matrix adaptation with Welford variances.  This kernel restates that
// algorithm so that, between two log-density evaluations, everything a chain has to do happens in ONE launch:
//
//     loop:  bplx_nuts_step  (finish the pending leapfrog, grow / merge trees, adapt, collect, start the next
//                             leapfrog: theta_eval = z + eps * M^-1 (r + eps/2 * grad))
//            bplx_logdensity_fwdbwd(theta_eval) -> lp, grad
//
// Chains do not run in lock step: a chain that finishes its tree starts its next transition in the same launch,
// so the number of launches is the longest chain's total leapfrog count, not the sum of per-draw maxima that a
// vmapped `while_loop` pays.  All vectors are chain-minor ([D][ld], lane = chain: every access is coalesced).
#include <curand_kernel.h>
#include <math_constants.h>

#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "nuts.h"

namespace bplx {

namespace {

constexpr int kNutsMaxY = 16;  // threads that share one chain (slices of the parameter axis)

struct Vec {  // one per-chain vector of length D inside a chain-minor array
  float* p;
  size_t ld;
  __device__ __forceinline__ float& operator[](int d) const { return p[(size_t)d * ld]; }
};

__device__ __forceinline__ float log_add_exp(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == -CUDART_INF_F) return m;
  return m + log1pf(expf(-fabsf(a - b)));
}

// chain-major state: the 32 lanes of a warp are the slices of ONE chain -- a butterfly gives every lane the same sum
template <int N>
__device__ __forceinline__ void reduce_lanes(float (&v)[N]) {
#pragma unroll
  for (int i = 0; i < N; i++) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], m);
  }
}

// sum over the threadIdx.y slices of every chain of the block (all threads of the block must call it)
template <int N>
__device__ __forceinline__ void reduce_y(float (&v)[N], float* red) {
  const int Y = blockDim.y, x = threadIdx.x, y = threadIdx.y;
  __syncthreads();  // the previous use of `red` is over
#pragma unroll
  for (int i = 0; i < N; i++) red[(i * kNutsMaxY + y) * 32 + x] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; i++) {
    float s = 0.0f;
    for (int k = 0; k < Y; k++) s += red[(i * kNutsMaxY + k) * 32 + x];  // same order in every slice: identical results
    v[i] = s;
  }
}

// Streaming diagnostics (include/bplx_nuts.h): draw number k (0-based, post warm-up) of the chain / parameter at element
// offset o of a [D][ld] plane; `plane` = D * ld elements.
__device__ __forceinline__ void diag_collect(const bplx_nuts_params& P, int k, size_t o, size_t plane, float x) {
  const int L = P.diag_lags, N = P.num_samples, h = N / 2;
  float ref = x;
  if (k == 0) P.dg_ref[o] = x;
  else ref = P.dg_ref[o];
  const float v = x - ref;
  float* s = P.dg_sums + o;
  s[0] += v;
  s[plane] = fmaf(v, v, s[plane]);
  if (k < h) {
    s[2 * plane] += v;
    s[3 * plane] = fmaf(v, v, s[3 * plane]);
  }
  if (k >= N - h) {
    s[4 * plane] += v;
    s[5 * plane] = fmaf(v, v, s[5 * plane]);
  }
  const int nl = k < L ? k : L;
  for (int l = 1; l <= nl; l++) {
    const float prev = P.dg_ring[(size_t)((k - l) % L) * plane + o];
    float* q = P.dg_lag + (size_t)(l - 1) * plane + o;
    *q = fmaf(v, prev, *q);
  }
  P.dg_ring[(size_t)(k % L) * plane + o] = v;
  if (k < L) P.dg_head[(size_t)k * plane + o] = v;
}

}  // namespace

// The generic kernel's loops over a thread's parameters d = y, y + Y, ... come in batches of kNutsBatch: every load of a
// batch is issued before the first store, so a thread has kNutsBatch x (vectors read) requests in flight instead of one
// round trip to L2 / HBM per parameter (the vectors may alias as far as the compiler knows, so it does not hoist the loads
// itself; measured on configs[2], 32,768 chains x 1,339 parameters: the step was 4x the log-density kernel).  The
// arithmetic and its order are unchanged: same bits as before, and as the register-resident kernel.
// (chain-major, measured on configs[2] at 32,768 chains: 4 blocks of 8 warps per SM at 64 registers -- cold chain-state
//  fields spill -- with batches of 4: 0.76 ms per step; batches of 2: 0.85, of 6: 0.82, of 8: 0.89; 2 blocks per SM at
//  128 registers: 1.07; 5 / 6 blocks per SM: 0.91 / 0.98)
constexpr int kNutsBatchCMinor = 4, kNutsBatchCMajor = 4;
#define BPLX_FOR_BATCH(d0) for (int d0 = y; d0 < D; d0 += kNutsBatch * Y)
#define BPLX_IN_BATCH(u, d, d0) \
  _Pragma("unroll") for (int u = 0, d = d0; u < kNutsBatch; u++, d += Y)

// block = (32 chains, Y slices).  Thread (x, y) owns parameters d = y, y+Y, ... of chain x; the scalar state machine
// is replicated in the Y threads of a chain (same inputs, same random numbers -> same decisions); only y == 0 writes it.
//
// CM = true (bplx_nuts_params::state_layout == 1, chain-major state): block = (32 lanes, chains), a WARP per chain; lane
// y owns d = y, y + 32, ... of the chain's contiguous vectors.  Chains are not in lock step, so in the chain-minor
// layout the lanes of a warp sit in different stages, extend different ends and test different checkpoint levels:
// every 128-byte line is touched for a few of its 32 chains, stores fill sectors partially (read-modify-write in L2)
// -- measured 11.2 GB of DRAM traffic per step on configs[2] against ~2.8 GB of vectors actually used.  With a warp
// per chain every branch is warp-uniform and every access a full line of one chain's vector; the reductions over the
// parameters are warp butterflies and the block never synchronises.
template <bool CM>
__global__ void __launch_bounds__(CM ? 256 : 32 * kNutsMaxY, CM ? 4 : 1) nuts_step_kernel(const bplx_nuts_params P) {
  constexpr int kNutsBatch = CM ? kNutsBatchCMajor : kNutsBatchCMinor;
  __shared__ float red[CM ? 1 : 2 * kNutsMaxY * 32];
  pdl_wait();  // (launched with programmatic stream serialization: nothing here may run ahead of the log-density kernel)
  pdl_launch_dependents();
  const int Y = CM ? 32 : blockDim.y, y = CM ? threadIdx.x : threadIdx.y;
  const int c_raw = CM ? blockIdx.x * blockDim.y + threadIdx.y : blockIdx.x * 32 + threadIdx.x;
  const bool valid = c_raw < P.C;
  const int c = valid ? c_raw : P.C - 1;
  const int D = P.D;
  // element (k, d) of chain c in a stack of vectors: chain-minor [k][D][ld] + c, chain-major [k][C][ld_state] + d
  const size_t ld = CM ? (size_t)1 : (size_t)P.ld;
  const size_t plane = CM ? (size_t)P.C * (size_t)P.ld_state : (size_t)D * (size_t)P.ld;
  const size_t cbase = CM ? (size_t)c * (size_t)P.ld_state : (size_t)c;
  auto vec = [&](float* base) { return Vec{base + cbase, ld}; };
  auto vec_k = [&](float* base, int k) { return Vec{base + (size_t)k * plane + cbase, ld}; };
  auto reduce = [&](auto& v) {
    if (CM) reduce_lanes(v);
    else reduce_y(v, red);
  };
  auto block_or = [&](bool x) { return CM ? x : (bool)__syncthreads_or(x); };  // (chain-major: a chain's warp decides alone)
  const Vec th = vec(P.theta_eval), gr = vec(P.grad), ph = vec(P.p_half), imm = vec(P.inv_mass);
  const Vec zL = vec(P.zL), rL = vec(P.rL), gL = vec(P.gL), zR = vec(P.zR), rR = vec(P.rR), gR = vec(P.gR);
  const Vec zP = vec(P.zP), gP = vec(P.gP), rS = vec(P.r_sum);
  const Vec zQ = vec(P.zQ), gQ = vec(P.gQ), rSq = vec(P.r_sum_sub);
  NutsChain* chains = static_cast<NutsChain*>(P.chain);
  NutsChain st = chains[c];
  const bool live = valid && st.stage != kNutsDone;  // threads of finished / padding chains only take part in barriers
  curandStatePhilox4_32_10_t rng;
  curand_init(P.seed, (unsigned long long)(P.chain_offset + c), st.rng_offset, &rng);
  unsigned draws = 0;
  auto uniform = [&]() { draws++; return curand_uniform(&rng); };
  const float lp_new = P.lp[c];

  // ======== A. finish the pending leapfrog: r1 = r_half + eps/2 grad(lp); the new leaf replaces the outer leaf ========
  const bool pending = live && st.stage == kNutsEvalPending;
  if (live && st.stage == kNutsInitEval) {  // gradient at the initial position has just been computed
    for (int d = y; d < D; d += Y) {
      zP[d] = th[d];
      gP[d] = gr[d];
    }
    st.pe = -lp_new;
    st.stage = kNutsNewTransition;
  }
  const float eps = st.going_right ? st.step_size : -st.step_size;
  const Vec zE = st.going_right ? zR : zL, rE = st.going_right ? rR : rL, gE = st.going_right ? gR : gL;
  // checkpoints for the iterative U-turn test (numpyro `_leaf_idx_to_ckpt_idxs`, `_is_iterative_turning`): the levels
  // idx_min .. idx_max this leaf is tested against, and the level an even leaf is stored at
  const unsigned leaf = (unsigned)st.sub_num;
  int idx_min = 1, idx_max = 0;
  if (pending) {
    idx_max = __popc(leaf >> 1);
    idx_min = idx_max - (__ffs(~leaf) - 1) + 1;  // minus the number of trailing one bits
  }
  const int nlv = (pending && idx_max >= idx_min) ? idx_max - idx_min + 1 : 0;  // levels idx_min .. idx_max
  const bool ckpt = pending && (leaf & 1u) == 0u;
  // ONE pass over the parameters does everything that does not depend on the accept / take decisions: the new leaf
  // (momentum, position, gradient of the extended end), the kinetic energy, the subtree's momentum sum, the checkpoint
  // an even leaf leaves behind and the U-turn dot products of every level the leaf is tested against -- each vector is
  // read once (the stage-by-stage form re-read the leaf's momentum, the sum and the inverse mass once per level).
  // (the dot products of kSlots levels at a time live in registers: a leaf is tested against as many levels as its index
  //  has trailing one bits plus one -- more than four for one leaf in sixteen, which then takes another pass below)
  constexpr int kSlots = 4;
  bool spec = false;
  float acc1[1] = {0.0f};
  float dt0[kSlots], dt1[kSlots];
#pragma unroll
  for (int j = 0; j < kSlots; j++) dt0[j] = dt1[j] = 0.0f;
  if (pending) {
    const Vec ckw = vec_k(P.r_ckpts, idx_max), cksw = vec_k(P.r_sum_ckpts, idx_max);
    const bool first = st.sub_num == 0;
    // a leaf that is not the last of its subtree is followed by a leapfrog from the same end, unless the subtree turns or
    // diverges: its start (stage F) is computed here, from registers, and redone below in the rare other case
    spec = st.sub_num + 1 < (1 << st.depth);
    BPLX_FOR_BATCH(d0) {
      float g[kNutsBatch], p[kNutsBatch], m[kNutsBatch], t[kNutsBatch], q[kNutsBatch], r1[kNutsBatch], rs[kNutsBatch];
      BPLX_IN_BATCH(u, d, d0) {
        const bool ok = d < D;
        g[u] = ok ? gr[d] : 0.0f;
        p[u] = ok ? ph[d] : 0.0f;
        m[u] = ok ? imm[d] : 0.0f;
        t[u] = ok ? th[d] : 0.0f;
        q[u] = (ok && !first) ? rSq[d] : 0.0f;
      }
      BPLX_IN_BATCH(u, d, d0) {
        r1[u] = fmaf(0.5f * eps, g[u], p[u]);
        rs[u] = first ? r1[u] : q[u] + r1[u];
      }
#pragma unroll
      for (int j = 0; j < kSlots; j++) {
        if (j < nlv) {
          const Vec ck = vec_k(P.r_ckpts, idx_min + j), cks = vec_k(P.r_sum_ckpts, idx_min + j);
          float a[kNutsBatch], k2[kNutsBatch];
          BPLX_IN_BATCH(u, d, d0) {
            const bool ok = d < D;
            a[u] = ok ? ck[d] : 0.0f;
            k2[u] = ok ? cks[d] : 0.0f;
          }
          BPLX_IN_BATCH(u, d, d0) {
            if (d < D) {
              const float sm = (rs[u] - k2[u] + a[u]) - 0.5f * (a[u] + r1[u]);  // momentum sum of the subtree that starts at the checkpoint
              dt0[j] = fmaf(m[u] * a[u], sm, dt0[j]);
              dt1[j] = fmaf(m[u] * r1[u], sm, dt1[j]);
            }
          }
        }
      }
      BPLX_IN_BATCH(u, d, d0) {
        if (d < D) {
          acc1[0] = fmaf(m[u] * r1[u], r1[u], acc1[0]);
          rE[d] = r1[u];
          zE[d] = t[u];
          gE[d] = g[u];
          rSq[d] = rs[u];
          if (ckpt) {
            ckw[d] = r1[u];
            cksw[d] = rs[u];
          }
          if (spec) {
            const float rh = fmaf(0.5f * eps, g[u], r1[u]);
            ph[d] = rh;
            th[d] = fmaf(eps * m[u], rh, t[u]);
          }
        }
      }
    }
  }
  reduce(acc1);
  bool sub_done = false;
  if (pending) {
    const float pe1 = -lp_new;
    float delta = pe1 + 0.5f * acc1[0] - st.energy_current;
    if (!(delta == delta) || !(fabsf(lp_new) < CUDART_INF_F)) delta = CUDART_INF_F;  // NaN, or a log-density that is not finite -> reject
    const float w_leaf = -delta;
    const float acc = fminf(1.0f, expf(-delta));
    // ---- merge the leaf into the subtree (uniform transition kernel inside a subtree) -----------------------------
    bool take;
    if (st.sub_num == 0) {
      take = true;
      st.sub_weight = w_leaf;
    } else {
      const float pr = 1.0f / (1.0f + expf(-(w_leaf - st.sub_weight)));
      take = uniform() < pr;
      st.sub_weight = log_add_exp(st.sub_weight, w_leaf);
    }
    if (take) {  // the subtree's proposal moves to the new leaf (its position: the copy just made, theta may have moved on)
      BPLX_FOR_BATCH(d0) {
        float t[kNutsBatch], g[kNutsBatch];
        BPLX_IN_BATCH(u, d, d0) {
          const bool ok = d < D;
          t[u] = ok ? zE[d] : 0.0f;
          g[u] = ok ? gr[d] : 0.0f;
        }
        BPLX_IN_BATCH(u, d, d0) {
          if (d < D) {
            zQ[d] = t[u];
            gQ[d] = g[u];
          }
        }
      }
      st.sub_pe = pe1;
    }
    st.sub_div = delta > P.max_delta_energy;
    st.sub_sum_accept += acc;
  }
  // ======== B. iterative U-turn test against the checkpoints (levels idx_max .. idx_min, stop at the first turn) =======
  // (the subtree turns if it turns against ANY of its levels: the order of the tests does not matter)
  bool turning = false;
  for (int j0 = 0; block_or(j0 < nlv && !turning); j0 += kSlots) {
    if (j0 > 0) {  // the next kSlots levels of a leaf with many: the new leaf's momentum and the sum are read back
#pragma unroll
      for (int j = 0; j < kSlots; j++) dt0[j] = dt1[j] = 0.0f;
      if (j0 < nlv && !turning) {
        BPLX_FOR_BATCH(d0) {
          float m[kNutsBatch], r1[kNutsBatch], rs[kNutsBatch];
          BPLX_IN_BATCH(u, d, d0) {
            const bool ok = d < D;
            m[u] = ok ? imm[d] : 0.0f;
            r1[u] = ok ? rE[d] : 0.0f;
            rs[u] = ok ? rSq[d] : 0.0f;
          }
#pragma unroll
          for (int j = 0; j < kSlots; j++) {
            if (j0 + j < nlv) {
              const Vec ck = vec_k(P.r_ckpts, idx_min + j0 + j), cks = vec_k(P.r_sum_ckpts, idx_min + j0 + j);
              float a[kNutsBatch], k2[kNutsBatch];
              BPLX_IN_BATCH(u, d, d0) {
                const bool ok = d < D;
                a[u] = ok ? ck[d] : 0.0f;
                k2[u] = ok ? cks[d] : 0.0f;
              }
              BPLX_IN_BATCH(u, d, d0) {
                if (d < D) {
                  const float sm = (rs[u] - k2[u] + a[u]) - 0.5f * (a[u] + r1[u]);
                  dt0[j] = fmaf(m[u] * a[u], sm, dt0[j]);
                  dt1[j] = fmaf(m[u] * r1[u], sm, dt1[j]);
                }
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kSlots; j++) {
      const bool need = j0 + j < nlv && !turning;
      if (!block_or(need)) continue;
      float dots[2] = {dt0[j], dt1[j]};
      reduce(dots);
      if (need) turning = dots[0] <= 0.0f || dots[1] <= 0.0f;
    }
  }
  // ======== C. subtree complete: merge into the trajectory (biased progressive sampling, numpyro `_combine_tree`) =======
  bool move = false;
  if (pending) {
    st.sub_turning = turning;
    st.sub_num += 1;
    st.stage = kNutsInTree;
    sub_done = st.sub_num == (1 << st.depth) || st.sub_turning || st.sub_div;
    if (sub_done) {
      float pr = fminf(1.0f, expf(st.sub_weight - st.weight));
      if (st.sub_turning || st.sub_div) pr = 0.0f;
      move = uniform() < pr;
    }
  }
  {
    float dots[2] = {0.0f, 0.0f};  // generalised U-turn of the whole trajectory (numpyro `_is_turning`, diagonal mass)
    if (sub_done) {
      BPLX_FOR_BATCH(d0) {
        float r0[kNutsBatch], q[kNutsBatch], zq[kNutsBatch], gq[kNutsBatch], a[kNutsBatch], b[kNutsBatch], m[kNutsBatch];
        BPLX_IN_BATCH(u, d, d0) {
          const bool ok = d < D;
          r0[u] = ok ? rS[d] : 0.0f;
          q[u] = ok ? rSq[d] : 0.0f;
          zq[u] = (ok && move) ? zQ[d] : 0.0f;
          gq[u] = (ok && move) ? gQ[d] : 0.0f;
          a[u] = ok ? rL[d] : 0.0f;
          b[u] = ok ? rR[d] : 0.0f;
          m[u] = ok ? imm[d] : 0.0f;
        }
        BPLX_IN_BATCH(u, d, d0) {
          if (d < D) {
            const float rs = r0[u] + q[u];
            rS[d] = rs;
            if (move) {
              zP[d] = zq[u];
              gP[d] = gq[u];
            }
            const float s = rs - 0.5f * (a[u] + b[u]);
            dots[0] = fmaf(m[u] * a[u], s, dots[0]);
            dots[1] = fmaf(m[u] * b[u], s, dots[1]);
          }
        }
      }
    }
    if (block_or(sub_done)) reduce(dots);
    if (sub_done) {
      if (move) st.pe = st.sub_pe;
      st.turning = st.sub_turning || dots[0] <= 0.0f || dots[1] <= 0.0f;
      st.depth += 1;
      st.weight = log_add_exp(st.weight, st.sub_weight);
      st.diverging = st.sub_div;
      st.sum_accept += st.sub_sum_accept;
      st.num_prop += st.sub_num;
      st.sub_active = 0;
    }
  }
  // ======== D. transition complete: adaptation during warm-up, collection afterwards ===============================
  if (sub_done && (st.depth >= P.max_tree_depth || st.turning || st.diverging)) {
    const float accept = st.sum_accept / (float)st.num_prop;
    st.num_leapfrog_total += st.num_prop;
    if (st.diverging && st.t >= P.num_warmup) st.num_divergent += 1;
    if (st.t < P.num_warmup) {
      // dual averaging (numpyro `dual_averaging`) on g = target - accept
      const float g = P.target_accept - accept;
      st.da_t += 1;
      const float t = (float)st.da_t;
      st.da_g_avg = (1.0f - 1.0f / (t + 10.0f)) * st.da_g_avg + g / (t + 10.0f);
      st.da_x = st.da_mu - sqrtf(t) / 0.05f * st.da_g_avg;
      const float wt = powf(t, -0.75f);
      st.da_x_avg = (1.0f - wt) * st.da_x_avg + wt * st.da_x;
      st.step_size = fmaxf(expf(st.t == P.num_warmup - 1 ? st.da_x_avg : st.da_x), 1.1754944e-38f);
      const bplx_window win = P.windows[st.window];  // the current adaptation window
      const bool middle = st.window > 0 && st.window < P.num_windows - 1;
      const bool at_end = st.t == win.end;
      if (middle) {  // Welford on the new position; at the window's end: regularised variance -> inverse mass
        const Vec mean = vec(P.wf_mean), m2 = vec(P.wf_m2);
        st.wf_n += 1;
        const float n = (float)st.wf_n, inv_n = 1.0f / n;
        for (int d = y; d < D; d += Y) {
          const float x = zP[d], pre = x - mean[d];
          const float mnew = fmaf(pre, inv_n, mean[d]);
          const float m2new = fmaf(pre, x - mnew, m2[d]);
          if (at_end) {
            imm[d] = (n / (n + 5.0f)) * (m2new / (n - 1.0f)) + 1e-3f * (5.0f / (n + 5.0f));
            mean[d] = 0.0f;
            m2[d] = 0.0f;
          } else {
            mean[d] = mnew;
            m2[d] = m2new;
          }
        }
        if (at_end) {  // restart the step-size search around 10 eps
          st.wf_n = 0;
          st.da_mu = logf(10.0f * st.step_size);
          st.da_x = st.da_x_avg = st.da_g_avg = 0.0f;
          st.da_t = 0;
        }
      }
      if (at_end) st.window += 1;
    } else {
      const int k = st.t - P.num_warmup;
      if (P.diag_lags > 0)
        for (int d = y; d < D; d += Y) diag_collect(P, k, (size_t)d * ld + cbase, plane, zP[d]);
      if (k % P.thin == 0 && k / P.thin < P.num_keep) {
        const int slot = k / P.thin;
        float* out = P.samples + (size_t)slot * plane + cbase;
        for (int d = y; d < D; d += Y) out[(size_t)d * ld] = zP[d];
        if (y == 0) {
          P.sample_lp[(size_t)slot * P.ld + c] = -st.pe;
          P.sample_accept[(size_t)slot * P.ld + c] = accept;
        }
      }
    }
    st.t += 1;
    st.stage = st.t >= P.num_warmup + P.num_samples ? kNutsDone : kNutsNewTransition;
  }
  // ======== E. momentum refresh: r ~ N(0, M), M = 1 / inv_mass ==========================================================
  const bool fresh = live && st.stage == kNutsNewTransition;
  {
    float ke[1] = {0.0f};
    if (fresh) {
      // normals come from their own stretch of the chain's Philox sequence, addressed by d: every slice sees the same r[d]
      const unsigned long long base = st.rng_offset + 64ull;
      for (int d = y; d < D; d += Y) {
        curandStatePhilox4_32_10_t rn;
        curand_init(P.seed, (unsigned long long)(P.chain_offset + c), base + 4ull * (unsigned long long)(d >> 1), &rn);
        const float2 n2 = curand_normal2(&rn);
        const float m = imm[d];
        const float r = ((d & 1) ? n2.y : n2.x) * rsqrtf(m);
        ke[0] = fmaf(m * r, r, ke[0]);
        rL[d] = r;
        rR[d] = r;
        rS[d] = r;
        const float z = zP[d], g = gP[d];
        zL[d] = z;
        zR[d] = z;
        gL[d] = g;
        gR[d] = g;
      }
    }
    if (block_or(fresh)) reduce(ke);
    if (fresh) {
      draws += 64u + 2u * (unsigned)D + 8u;
      st.energy_current = st.pe + 0.5f * ke[0];
      st.depth = 0;
      st.weight = 0.0f;
      st.turning = st.diverging = 0;
      st.sum_accept = 0.0f;
      st.num_prop = 0;
      st.sub_active = 0;
      st.stage = kNutsInTree;
    }
  }
  // ======== F. start the next leapfrog from the outer leaf: r_half = r + eps/2 grad(lp); theta = z + eps M^-1 r_half ====
  if (live && st.stage == kNutsInTree) {
    if (!st.sub_active) {  // next doubling: direction, empty subtree
      st.going_right = uniform() < 0.5f ? 1 : 0;
      st.sub_active = 1;
      st.sub_num = 0;
      st.sub_weight = 0.0f;
      st.sub_sum_accept = 0.0f;
      st.sub_turning = st.sub_div = 0;
    }
    const float e2 = st.going_right ? st.step_size : -st.step_size;
    const Vec zF = st.going_right ? zR : zL, rF = st.going_right ? rR : rL, gF = st.going_right ? gR : gL;
    if (!(spec && !sub_done))  // (else: computed in the pass above, same arithmetic)
    BPLX_FOR_BATCH(d0) {
      float g[kNutsBatch], r0[kNutsBatch], m[kNutsBatch], z[kNutsBatch];
      BPLX_IN_BATCH(u, d, d0) {
        const bool ok = d < D;
        g[u] = ok ? gF[d] : 0.0f;
        r0[u] = ok ? rF[d] : 0.0f;
        m[u] = ok ? imm[d] : 0.0f;
        z[u] = ok ? zF[d] : 0.0f;
      }
      BPLX_IN_BATCH(u, d, d0) {
        if (d < D) {
          const float rh = fmaf(0.5f * e2, g[u], r0[u]);
          ph[d] = rh;
          th[d] = fmaf(e2 * m[u], rh, z[u]);
        }
      }
    }
    st.stage = kNutsEvalPending;
  }
  if (live && y == 0) {
    st.rng_offset += (draws + 7u) & ~3u;  // curand_init's offset counts 32-bit outputs; keep calls on disjoint ranges
    chains[c] = st;
    if (st.stage != kNutsDone) atomicAdd(P.active_count, 1);
  }
}

// ---- register-resident variant -------------------------------------------------------------------------------------
// The kernel above walks its stages through global memory: every stage re-reads vectors the previous one wrote, so a
// launch is a chain of dependent round trips to L2 -- and with many chains per block the block pays for the union of
// the stages its chains are in (measured, DixonColes D = 44: 10 us at one chain, 46 us at 1,024 chains, three times the
// log-density kernel beside it).  For small models the slice of every state vector a thread owns fits in registers:
// block = (CPB chains, Y = 512 / CPB slices), thread (x, y) owns d = y + j Y, j < NPT.  Everything that does not
// depend on the chain's scalar state is requested in ONE batch at entry; what does (the other end of the trajectory,
// the Welford accumulators, the first checkpoint level of the U-turn test) in a second batch that arrives while the
// first stages run; the stages then work on registers and store what they change.  Same arithmetic in the same order
// as the kernel above: with the same (32, Y) geometry the two give identical bits (tests/test_nuts_gpu.py).
// CM = true (chain-major state, bplx_nuts_params::state_layout == 1): block = (32 lanes, CPB chains), a warp per chain,
// lane y owns d = y + 32 j -- the reductions are warp butterflies, the U-turn levels a chain's own, no block barrier.
template <int CPB, int NPT, bool CM>
__global__ void __launch_bounds__(CM ? 32 * CPB : 512, CM ? 2 : 1) nuts_step_fast_kernel(const bplx_nuts_params P) {
  __shared__ float red[CM ? 1 : 2 * 512];
  __shared__ unsigned lvl_mask;
  const int Y = CM ? 32 : blockDim.y, y = CM ? threadIdx.x : threadIdx.y, x = CM ? threadIdx.y : threadIdx.x;
  const int c_raw = blockIdx.x * CPB + x;
  const bool valid = c_raw < P.C;
  const int c = valid ? c_raw : P.C - 1;
  const int D = P.D;
  const size_t ld = CM ? (size_t)1 : (size_t)P.ld;  // stride between a chain's consecutive parameters
  const size_t plane = CM ? (size_t)P.C * (size_t)P.ld_state : (size_t)D * (size_t)P.ld;  // one vector of every chain
  const size_t cbase = CM ? (size_t)c * (size_t)P.ld_state : (size_t)c;
  if (!CM && x == 0 && y == 0) lvl_mask = 0u;
  bool has[NPT];
  size_t off[NPT];
#pragma unroll
  for (int j = 0; j < NPT; j++) {
    const int d = y + j * Y;
    has[j] = d < D;
    off[j] = (size_t)(has[j] ? d : 0) * ld + cbase;
  }
  auto ldv = [&](const float* base, float (&v)[NPT], bool on) {
#pragma unroll
    for (int j = 0; j < NPT; j++) v[j] = (on && has[j]) ? base[off[j]] : 0.0f;
  };
  auto stv = [&](float* base, const float (&v)[NPT]) {
#pragma unroll
    for (int j = 0; j < NPT; j++)
      if (has[j]) base[off[j]] = v[j];
  };
  auto reduce1 = [&](float& v) {  // sum over the slices of every chain (all threads of the block call it)
    if (CM) {
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
      return;
    }
    __syncthreads();
    red[y * CPB + x] = v;
    __syncthreads();
    float s = 0.0f;
    for (int k = 0; k < Y; k++) s += red[k * CPB + x];
    v = s;
  };
  auto reduce2 = [&](float (&v)[2]) {
    if (CM) {
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], m);
        v[1] += __shfl_xor_sync(0xffffffffu, v[1], m);
      }
      return;
    }
    __syncthreads();
    red[y * CPB + x] = v[0];
    red[(Y + y) * CPB + x] = v[1];
    __syncthreads();
    float s0 = 0.0f, s1 = 0.0f;
    for (int k = 0; k < Y; k++) {
      s0 += red[k * CPB + x];
      s1 += red[(Y + k) * CPB + x];
    }
    v[0] = s0;
    v[1] = s1;
  };

  // ---- batch 1: independent of the chain's state ---------------------------------------------------------------------
  // (under programmatic dependent launch this batch, the chain state and batch 2 overlap the tail of the log-density
  //  kernel: it writes none of them; its outputs -- lp, grad -- are read after pdl_wait() below)
  float th[NPT], gr[NPT], ph[NPT], im[NPT], rL[NPT], rR[NPT], zP[NPT], gP[NPT], rS[NPT], zQ[NPT], gQ[NPT], rQ[NPT];
  ldv(P.theta_eval, th, true);
  ldv(P.p_half, ph, true);
  ldv(P.inv_mass, im, true);
  ldv(P.rL, rL, true);
  ldv(P.rR, rR, true);
  ldv(P.zP, zP, true);
  ldv(P.gP, gP, true);
  ldv(P.r_sum, rS, true);
  ldv(P.zQ, zQ, true);
  ldv(P.gQ, gQ, true);
  ldv(P.r_sum_sub, rQ, true);
  NutsChain* chains = static_cast<NutsChain*>(P.chain);
  NutsChain st = chains[c];
  const bool live = valid && st.stage != kNutsDone;
  const bool pending = live && st.stage == kNutsEvalPending;
  const bool right = st.going_right != 0;  // the direction of the leapfrog that has just been evaluated
  // the U-turn levels this leaf has to be tested against (numpyro `_leaf_idx_to_ckpt_idxs`)
  const unsigned leaf = (unsigned)st.sub_num;
  int idx_min = 1, idx_max = 0;
  if (pending) {
    idx_max = __popc(leaf >> 1);
    idx_min = idx_max - (__ffs(~leaf) - 1) + 1;  // minus the number of trailing one bits
  }
  const unsigned my_lvls = (pending && idx_max >= idx_min) ? (((2u << idx_max) - 1u) & ~((1u << idx_min) - 1u)) : 0u;
  // ---- batch 2: the other end of the trajectory, the Welford accumulators, the first checkpoint level ----------------
  float zO[NPT], gO[NPT], wm[NPT], w2[NPT], ck[NPT], cks[NPT];
  ldv(right ? P.zL : P.zR, zO, live);
  ldv(right ? P.gL : P.gR, gO, live);
  const bool warm = live && st.t < P.num_warmup;
  ldv(P.wf_mean, wm, warm);
  ldv(P.wf_m2, w2, warm);
  {
    const int i0 = my_lvls ? 31 - __clz(my_lvls) : 0;
    ldv(P.r_ckpts + (size_t)i0 * plane, ck, my_lvls != 0u);
    ldv(P.r_sum_ckpts + (size_t)i0 * plane, cks, my_lvls != 0u);
  }
  curandStatePhilox4_32_10_t rng;
  curand_init(P.seed, (unsigned long long)(P.chain_offset + c), st.rng_offset, &rng);
  unsigned draws = 0;
  auto uniform = [&]() { draws++; return curand_uniform(&rng); };
  // ---- the log-density kernel's outputs ---------------------------------------------------------------------------------
  pdl_wait();
  pdl_launch_dependents();
  ldv(after_wait(P.grad), gr, true);
  const float lp_new = after_wait(P.lp)[c];

  // ======== A. finish the pending leapfrog ==============================================================================
  if (live && st.stage == kNutsInitEval) {  // gradient at the initial position has just been computed
#pragma unroll
    for (int j = 0; j < NPT; j++) {
      zP[j] = th[j];
      gP[j] = gr[j];
    }
    stv(P.zP, zP);
    stv(P.gP, gP);
    st.pe = -lp_new;
    st.stage = kNutsNewTransition;
  }
  const float eps = right ? st.step_size : -st.step_size;
  float r1[NPT];
  float acc1 = 0.0f;
#pragma unroll
  for (int j = 0; j < NPT; j++) r1[j] = 0.0f;
  if (pending) {
#pragma unroll
    for (int j = 0; j < NPT; j++) {
      if (has[j]) {
        r1[j] = fmaf(0.5f * eps, gr[j], ph[j]);
        acc1 = fmaf(im[j] * r1[j], r1[j], acc1);
        if (right) rR[j] = r1[j];
        else rL[j] = r1[j];
      }
    }
    stv(right ? P.rR : P.rL, r1);
    stv(right ? P.zR : P.zL, th);
    stv(right ? P.gR : P.gL, gr);
  }
  reduce1(acc1);
  bool sub_done = false;
  if (pending) {
    const float pe1 = -lp_new;
    float delta = pe1 + 0.5f * acc1 - st.energy_current;
    if (!(delta == delta) || !(fabsf(lp_new) < CUDART_INF_F)) delta = CUDART_INF_F;  // NaN, or a log-density that is not finite -> reject
    const float w_leaf = -delta;
    const float acc = fminf(1.0f, expf(-delta));
    bool take;
    if (st.sub_num == 0) {
      take = true;
      st.sub_weight = w_leaf;
    } else {
      const float pr = 1.0f / (1.0f + expf(-(w_leaf - st.sub_weight)));
      take = uniform() < pr;
      st.sub_weight = log_add_exp(st.sub_weight, w_leaf);
    }
    const bool ckpt = (leaf & 1u) == 0u;
#pragma unroll
    for (int j = 0; j < NPT; j++) {
      rQ[j] = st.sub_num == 0 ? r1[j] : rQ[j] + r1[j];
      if (take) {
        zQ[j] = th[j];
        gQ[j] = gr[j];
      }
    }
    stv(P.r_sum_sub, rQ);
    if (take) {
      stv(P.zQ, zQ);
      stv(P.gQ, gQ);
    }
    if (ckpt) {
      stv(P.r_ckpts + (size_t)idx_max * plane, r1);
      stv(P.r_sum_ckpts + (size_t)idx_max * plane, rQ);
    }
    if (take) st.sub_pe = pe1;
    st.sub_div = delta > P.max_delta_energy;
    st.sub_sum_accept += acc;
  }
  // ======== B. iterative U-turn test against the checkpoints: only the levels some chain of the block needs ============
  if (!CM) {
    if (my_lvls && y == 0) atomicOr(&lvl_mask, my_lvls);
    __syncthreads();
  }
  bool turning = false;
  {
    unsigned m = CM ? my_lvls : lvl_mask;
    // (the level prefetched in batch 2 is this chain's highest; the block's may be higher)
    int have = my_lvls ? 31 - __clz(my_lvls) : -1;
    while (m) {
      const int i = 31 - __clz(m);
      m &= ~(1u << i);
      const bool need = pending && ((my_lvls >> i) & 1u) && !turning;
      if (need && have != i) {
        ldv(P.r_ckpts + (size_t)i * plane, ck, true);
        ldv(P.r_sum_ckpts + (size_t)i * plane, cks, true);
        have = i;
      }
      float dots[2] = {0.0f, 0.0f};
      if (need) {
#pragma unroll
        for (int j = 0; j < NPT; j++) {
          if (has[j]) {
            const float a = ck[j], b = right ? rR[j] : rL[j], mm = im[j];
            const float sm = (rQ[j] - cks[j] + a) - 0.5f * (a + b);  // momentum sum of the subtree that starts at checkpoint i
            dots[0] = fmaf(mm * a, sm, dots[0]);
            dots[1] = fmaf(mm * b, sm, dots[1]);
          }
        }
      }
      // the next level this chain needs is requested before the reduction's barriers
      const unsigned below = my_lvls & ((1u << i) - 1u);
      if (need && below) {
        const int inext = 31 - __clz(below);
        ldv(P.r_ckpts + (size_t)inext * plane, ck, true);
        ldv(P.r_sum_ckpts + (size_t)inext * plane, cks, true);
        have = inext;
      }
      reduce2(dots);
      if (need) turning = dots[0] <= 0.0f || dots[1] <= 0.0f;
    }
  }
  // ======== C. subtree complete: merge into the trajectory ===============================================================
  bool move = false;
  if (pending) {
    st.sub_turning = turning;
    st.sub_num += 1;
    st.stage = kNutsInTree;
    sub_done = st.sub_num == (1 << st.depth) || st.sub_turning || st.sub_div;
    if (sub_done) {
      float pr = fminf(1.0f, expf(st.sub_weight - st.weight));
      if (st.sub_turning || st.sub_div) pr = 0.0f;
      move = uniform() < pr;
    }
  }
  {
    float dots[2] = {0.0f, 0.0f};
    if (sub_done) {
#pragma unroll
      for (int j = 0; j < NPT; j++) {
        if (has[j]) {
          const float rs = rS[j] + rQ[j];
          rS[j] = rs;
          if (move) {
            zP[j] = zQ[j];
            gP[j] = gQ[j];
          }
          const float a = rL[j], b = rR[j], mm = im[j];
          const float sm = rs - 0.5f * (a + b);
          dots[0] = fmaf(mm * a, sm, dots[0]);
          dots[1] = fmaf(mm * b, sm, dots[1]);
        }
      }
      stv(P.r_sum, rS);
      if (move) {
        stv(P.zP, zP);
        stv(P.gP, gP);
      }
    }
    if (CM ? sub_done : (bool)__syncthreads_or(sub_done)) reduce2(dots);
    if (sub_done) {
      if (move) st.pe = st.sub_pe;
      st.turning = st.sub_turning || dots[0] <= 0.0f || dots[1] <= 0.0f;
      st.depth += 1;
      st.weight = log_add_exp(st.weight, st.sub_weight);
      st.diverging = st.sub_div;
      st.sum_accept += st.sub_sum_accept;
      st.num_prop += st.sub_num;
      st.sub_active = 0;
    }
  }
  // ======== D. transition complete: adaptation during warm-up, collection afterwards ===============================
  if (sub_done && (st.depth >= P.max_tree_depth || st.turning || st.diverging)) {
    const float accept = st.sum_accept / (float)st.num_prop;
    st.num_leapfrog_total += st.num_prop;
    if (st.diverging && st.t >= P.num_warmup) st.num_divergent += 1;
    if (st.t < P.num_warmup) {
      const float g = P.target_accept - accept;
      st.da_t += 1;
      const float t = (float)st.da_t;
      st.da_g_avg = (1.0f - 1.0f / (t + 10.0f)) * st.da_g_avg + g / (t + 10.0f);
      st.da_x = st.da_mu - sqrtf(t) / 0.05f * st.da_g_avg;
      const float wt = powf(t, -0.75f);
      st.da_x_avg = (1.0f - wt) * st.da_x_avg + wt * st.da_x;
      st.step_size = fmaxf(expf(st.t == P.num_warmup - 1 ? st.da_x_avg : st.da_x), 1.1754944e-38f);
      const bplx_window win = P.windows[st.window];
      const bool middle = st.window > 0 && st.window < P.num_windows - 1;
      const bool at_end = st.t == win.end;
      if (middle) {
        st.wf_n += 1;
        const float n = (float)st.wf_n, inv_n = 1.0f / n;
#pragma unroll
        for (int j = 0; j < NPT; j++) {
          const float xx = zP[j], pre = xx - wm[j];
          const float mnew = fmaf(pre, inv_n, wm[j]);
          const float m2new = fmaf(pre, xx - mnew, w2[j]);
          if (at_end) {
            if (has[j]) im[j] = (n / (n + 5.0f)) * (m2new / (n - 1.0f)) + 1e-3f * (5.0f / (n + 5.0f));
            wm[j] = 0.0f;
            w2[j] = 0.0f;
          } else {
            wm[j] = mnew;
            w2[j] = m2new;
          }
        }
        stv(P.wf_mean, wm);
        stv(P.wf_m2, w2);
        if (at_end) {
          stv(P.inv_mass, im);
          st.wf_n = 0;
          st.da_mu = logf(10.0f * st.step_size);
          st.da_x = st.da_x_avg = st.da_g_avg = 0.0f;
          st.da_t = 0;
        }
      }
      if (at_end) st.window += 1;
    } else {
      const int k = st.t - P.num_warmup;
      if (P.diag_lags > 0) {
#pragma unroll
        for (int j = 0; j < NPT; j++)
          if (has[j]) diag_collect(P, k, off[j], plane, zP[j]);
      }
      if (k % P.thin == 0 && k / P.thin < P.num_keep) {
        const int slot = k / P.thin;
        stv(P.samples + (size_t)slot * plane, zP);
        if (y == 0) {
          P.sample_lp[(size_t)slot * P.ld + c] = -st.pe;
          P.sample_accept[(size_t)slot * P.ld + c] = accept;
        }
      }
    }
    st.t += 1;
    st.stage = st.t >= P.num_warmup + P.num_samples ? kNutsDone : kNutsNewTransition;
  }
  // ======== E. momentum refresh: r ~ N(0, M), M = 1 / inv_mass ==========================================================
  const bool fresh = live && st.stage == kNutsNewTransition;
  {
    float ke = 0.0f;
    if (fresh) {
      const unsigned long long base = st.rng_offset + 64ull;
#pragma unroll
      for (int j = 0; j < NPT; j++) {
        if (has[j]) {
          const int d = y + j * Y;
          curandStatePhilox4_32_10_t rn;
          curand_init(P.seed, (unsigned long long)(P.chain_offset + c), base + 4ull * (unsigned long long)(d >> 1), &rn);
          const float2 n2 = curand_normal2(&rn);
          const float mm = im[j];
          const float r = ((d & 1) ? n2.y : n2.x) * rsqrtf(mm);
          ke = fmaf(mm * r, r, ke);
          rL[j] = r;
          rR[j] = r;
          rS[j] = r;
        }
      }
      stv(P.rL, rL);
      stv(P.rR, rR);
      stv(P.r_sum, rS);
      stv(P.zL, zP);
      stv(P.zR, zP);
      stv(P.gL, gP);
      stv(P.gR, gP);
    }
    if (CM ? fresh : (bool)__syncthreads_or(fresh)) reduce1(ke);
    if (fresh) {
      draws += 64u + 2u * (unsigned)D + 8u;
      st.energy_current = st.pe + 0.5f * ke;
      st.depth = 0;
      st.weight = 0.0f;
      st.turning = st.diverging = 0;
      st.sum_accept = 0.0f;
      st.num_prop = 0;
      st.sub_active = 0;
      st.stage = kNutsInTree;
    }
  }
  // ======== F. start the next leapfrog from the outer leaf ================================================================
  if (live && st.stage == kNutsInTree) {
    if (!st.sub_active) {
      st.going_right = uniform() < 0.5f ? 1 : 0;
      st.sub_active = 1;
      st.sub_num = 0;
      st.sub_weight = 0.0f;
      st.sub_sum_accept = 0.0f;
      st.sub_turning = st.sub_div = 0;
    }
    const bool nr = st.going_right != 0;
    const float e2 = nr ? st.step_size : -st.step_size;
    // the outer leaf on that side: after a refresh both ends are the current position; the end that has just been
    // extended is the leaf evaluated in this launch; the other end came with batch 2
    const bool same = nr == right;
#pragma unroll
    for (int j = 0; j < NPT; j++) {
      const float zF = fresh ? zP[j] : (same ? th[j] : zO[j]);
      const float gF = fresh ? gP[j] : (same ? gr[j] : gO[j]);
      const float rF = nr ? rR[j] : rL[j];
      const float rh = fmaf(0.5f * e2, gF, rF);
      ph[j] = rh;
      th[j] = fmaf(e2 * im[j], rh, zF);
    }
    stv(P.p_half, ph);
    stv(P.theta_eval, th);
    st.stage = kNutsEvalPending;
  }
  if (live && y == 0) {
    st.rng_offset += (draws + 7u) & ~3u;
    chains[c] = st;
    if (st.stage != kNutsDone) atomicAdd(P.active_count, 1);
  }
}

__global__ void nuts_init_kernel(const bplx_nuts_params P) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.C) return;
  NutsChain st{};
  st.stage = kNutsInitEval;
  st.step_size = P.init_step_size;
  st.da_mu = logf(10.0f * P.init_step_size);
  st.rng_offset = 0;
  static_cast<NutsChain*>(P.chain)[c] = st;
  for (int d = 0; d < P.D; d++) {
    const size_t o = P.state_layout == 1 ? (size_t)c * P.ld_state + d : (size_t)d * P.ld + c;
    P.inv_mass[o] = 1.0f;
    P.wf_mean[o] = 0.0f;
    P.wf_m2[o] = 0.0f;
  }
}

}  // namespace bplx

using namespace bplx;

extern "C" {

size_t bplx_nuts_chain_bytes(void) { return sizeof(NutsChain); }

int bplx_nuts_init(const bplx_nuts_params* p, void* stream) {
  BPLX_REQUIRE(p && p->C > 0 && p->D > 0 && p->ld >= p->C, BPLX_E_INVALID, "nuts: bad C / D / ld");
  BPLX_REQUIRE(p->state_layout == 0 || (p->state_layout == 1 && p->ld_state >= p->D), BPLX_E_INVALID,
               "nuts: state_layout must be 0 (chain-minor) or 1 (chain-major, ld_state >= D)");
  nuts_init_kernel<<<(p->C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(*p);
  BPLX_CUDA(cudaGetLastError());
  note_launch(1);
  return BPLX_OK;
}

int bplx_nuts_summary(const bplx_nuts_params* p, float* out_host) {
  BPLX_REQUIRE(p && out_host && p->C > 0, BPLX_E_INVALID, "nuts summary: bad arguments");
  std::vector<NutsChain> h((size_t)p->C);
  BPLX_CUDA(cudaMemcpy(h.data(), p->chain, h.size() * sizeof(NutsChain), cudaMemcpyDeviceToHost));
  for (int c = 0; c < p->C; c++) {
    float* o = out_host + (size_t)c * 8;
    o[0] = (float)h[c].t;
    o[1] = h[c].step_size;
    o[2] = (float)h[c].num_divergent;
    o[3] = (float)h[c].num_leapfrog_total;
    o[4] = (float)h[c].depth;
    o[5] = o[6] = o[7] = 0.0f;
  }
  return BPLX_OK;
}

int bplx_nuts_step(const bplx_nuts_params* p, void* stream) {
  BPLX_REQUIRE(p && p->C > 0 && p->D > 0 && p->ld >= p->C, BPLX_E_INVALID, "nuts: bad C / D / ld");
  BPLX_REQUIRE(p->max_tree_depth >= 1 && p->max_tree_depth <= 12 && p->thin >= 1, BPLX_E_INVALID,
               "nuts: max_tree_depth must be in [1, 12] and thin >= 1");
  BPLX_REQUIRE(p->diag_lags >= 0 && (p->diag_lags == 0 || (p->dg_ref && p->dg_sums && p->dg_lag && p->dg_ring && p->dg_head)),
               BPLX_E_INVALID, "nuts: diag_lags > 0 needs the dg_* accumulators");
  BPLX_REQUIRE(p->state_layout == 0 || (p->state_layout == 1 && p->ld_state >= p->D), BPLX_E_INVALID,
               "nuts: state_layout must be 0 (chain-minor) or 1 (chain-major, ld_state >= D)");
  int Y = (p->D + 3) / 4;  // about four parameters per thread, at most kNutsMaxY slices per chain
  Y = Y < 1 ? 1 : (Y > kNutsMaxY ? kNutsMaxY : Y);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool generic = env_switches().nuts_generic;  // testing: the stage-by-stage kernel for every size
  cudaLaunchConfig_t cfg{};
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // the step's state loads overlap the log-density kernel's tail
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1u : 0u;
  auto launch = [&](auto* fn, int cpb, dim3 block) {
    cfg.gridDim = dim3((unsigned)((p->C + cpb - 1) / cpb));
    cfg.blockDim = block;
    return cudaLaunchKernelEx(&cfg, fn, *p);
  };
  if (p->state_layout == 1 && !generic && p->D <= 64)  // chain-major, small model: the chain's slices live in registers
    BPLX_CUDA(launch(&nuts_step_fast_kernel<8, 2, true>, 8, dim3(32, 8)));
  else if (p->state_layout == 1 && !generic && p->D <= 128)
    BPLX_CUDA(launch(&nuts_step_fast_kernel<8, 4, true>, 8, dim3(32, 8)));
  else if (p->state_layout == 1)  // chain-major state: a warp per chain, eight chains per block
    BPLX_CUDA(launch(&nuts_step_kernel<true>, 8, dim3(32, 8)));
  else if (!generic && p->D <= 4 * kNutsMaxY)  // same geometry as the generic kernel: identical bits
    BPLX_CUDA(launch(&nuts_step_fast_kernel<32, 4, false>, 32, dim3(32, Y)));
  else if (!generic && p->D <= 128)
    BPLX_CUDA(launch(&nuts_step_fast_kernel<16, 4, false>, 16, dim3(16, 32)));
  else if (!generic && p->D <= 256)
    BPLX_CUDA(launch(&nuts_step_fast_kernel<8, 4, false>, 8, dim3(8, 64)));
  else
    BPLX_CUDA(launch(&nuts_step_kernel<false>, 32, dim3(32, Y)));
  BPLX_CUDA(cudaGetLastError());
  note_launch(1);
  return BPLX_OK;
}

}  // extern "C"
