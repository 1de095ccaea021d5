// logdensity.cu -- K1: fused Dixon-Coles-family log-density + gradient, one lane per chain.
//
// Replaces value_and_grad(potential_fn) of the reference `_model`s (bpl/dixon_coles.py:39-84,
// bpl/extended_dixon_coles.py:78-248, bpl/neutral_dixon_coles.py:102-283,
// bpl/neutral_dixon_coles_WC.py:83-232) together with bpl/_util.py:17-93, for a batch of chains.
//
// CTA = 32 chains (lane = chain) x nwarps warps.  Flow (DESIGN.md "K1"):
//   prologue  each warp takes teams t = warp, warp+W, ...: reads theta and writes the rows (exp of
//             the six exponents) of its virtual teams into the shared tables; static sum(w y eta).
//   phase 1   each warp walks its stream of lists: acc += w * table[opp]; per list
//             sum_w_lambda = own * acc (both rates), running maxima of lambda_h, lambda_a,
//             lambda_h*lambda_a with the list they came from (bpl/_util.py:23-30).  A team's
//             gradient wrt its log-rate halves goes to its raw slots in the grad buffer.
//   bounds    maxima reduced over warps -> LB, UB, corr_coef; every warp scans a slice of each
//             chain's arg-max lists for the entry that attained the maximum.
//   phase 2   tau lists (0-0, 1-0, 0-1 matches): log tau, d/d eta (added to the raw slots),
//             d/d corr_coef.
//   fix-up    d corr_coef / d eta of the arg-max matches is added.
//   team pass raw slots -> parameter gradients: priors, chain rule, hyper-parameter sums.
//   epilogue  cross-warp reduction, hyper priors + Jacobians, covariate coefficients; lp, corr_coef.
// A raw slot is written by the one thread that owns (team, chain) in a phase and red.add'ed by one
// thread per later stage, in program order: the result is deterministic.
//
// Few chains (fewer groups of 32 than SMs): kp.split = 2, 4 or 8 CTAs form a thread-block cluster on ONE group of
// chains.  Every CTA builds all tables, the list streams are dealt over split * W virtual warps, the CTA-wide barriers
// become cluster barriers and the cross-warp reductions (maxima, arg-max search, d/d corr_coef, hyper sums) land in the
// shared memory of cluster rank 0 through distributed shared memory.
#include <cooperative_groups.h>

#include <math.h>

#include "k1_common.cuh"
#include "problem.h"

namespace cg = cooperative_groups;

#ifdef BPLX_TIMELINE
// debug build only: per-warp globaltimer stamps at the region boundaries of CTA 0 (read back with bplx_timeline_read)
__device__ unsigned long long g_timeline[32 * 24];
#define BPLX_STAMP(i)                                                                    \
  do {                                                                                   \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {                                    \
      unsigned long long t_;                                                             \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                             \
      g_timeline[(threadIdx.x >> 5) * 24 + (i)] = t_;                                    \
    }                                                                                    \
  } while (0)
// the stamp is taken once `dep` (a float) has been computed
#define BPLX_STAMP_DEP(i, dep)                                                           \
  do {                                                                                   \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {                                    \
      unsigned long long t_;                                                             \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_) : "f"(dep));                  \
      g_timeline[(threadIdx.x >> 5) * 24 + (i)] = t_;                                    \
    }                                                                                    \
  } while (0)
#else
#define BPLX_STAMP(i)
#define BPLX_STAMP_DEP(i, dep)
#endif

namespace bplx {

// One phase-1 piece of a model that clips its rates at 15 (DIXON_COLES, EXTENDED; bpl/dixon_coles.py:66-75): entries
// (opponent row, w, w y_x, w y_y).  *lp  sum w (y log rate - rate) of a home list;  m1..m3  maxima of the two rates and
// their product;  *gx, *gy  d/d (log X, log Y) of the list's own team.
// FAST: the prologue's bound says no rate of these 32 chains reaches the clip -- the same arithmetic with
// min(x, 15) = x and every guard true, bit for bit.
template <bool FAST, bool HOME>
__device__ __forceinline__ void rate_piece_clip(uint32_t& a, const uint32_t e_end, const float2 own, const uint32_t tab,
                                                float& lp, float& m1, float& m2, float& m3, float& gx, float& gy) {
  float2 lg = make_float2(0.0f, 0.0f), lw = lg, g = lg;
  walk16(a, e_end, [&](const uint32_t at) {
    const uint4 q = lds128u(at);
    const float2 ea = lds64(tab + q.x);
    const float w = __uint_as_float(q.y);
    const float2 wy = make_float2(__uint_as_float(q.z), __uint_as_float(q.w));
    const float2 XY = mul2(own, ea);
    const float2 d = fma2(bc2(-w), XY, wy);  // w (y - rate): the gradient term of an unclipped rate
    if (FAST) {
      g = add2(g, d);
    } else {
      if (XY.x < 15.0f) g.x += d.x;
      if (XY.y < 15.0f) g.y += d.y;
    }
    if (HOME) {
      const float2 c = FAST ? XY : make_float2(fminf(XY.x, 15.0f), fminf(XY.y, 15.0f));
      lg = fma2(wy, make_float2(lg2_approx(c.x), lg2_approx(c.y)), lg);
      lw = fma2(bc2(w), c, lw);
      m1 = fmaxf(m1, c.x);
      m2 = fmaxf(m2, c.y);
      m3 = fmaxf(m3, c.x * c.y);
    }
  });
  gx = g.x;
  gy = g.y;
  if (HOME) lp = fmaf(lg.x + lg.y, kLn2, -(lw.x + lw.y));
}

// order-preserving float <-> unsigned key (for an integer atomicMax over floats of either sign)
__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t b = __float_as_uint(f);
  return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float(k ^ ((k & 0x80000000u) ? 0x80000000u : 0xffffffffu));
}

template <bool CLIP>
__global__ void __launch_bounds__(kMaxWarps * 32, 1) logdensity_kernel(const __grid_constant__ KernelParams kp,
                                                                        const __grid_constant__ WarpBounds wb) {
  extern __shared__ __align__(1024) unsigned char smem[];
  BPLX_STAMP(19);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = kp.nwarps;
  const ThetaOffsets& o = kp.off;
  const int S = kp.split;  // CTAs of the cluster that shares this group of chains (1: no cluster)
  const int crank = S > 1 ? (int)cg::this_cluster().block_rank() : 0;
  const int group = kp.group0 + (int)blockIdx.x / S;
  const int vwarp = crank * W + warp, VW = S * W;  // virtual warp: owner of teams / streams across the cluster
  auto sync_all = [&]() {
    if (S > 1) cg::this_cluster().sync();
    else __syncthreads();
  };
  const int chain_raw = group * kChains + lane;
  const int chain = min(chain_raw, kp.C - 1);
  Lane ln;
  ln.th = kp.theta + (size_t)chain * (size_t)kp.sc;
  ln.gr = kp.grad + (size_t)chain * (size_t)kp.sc;
  ln.sc = kp.scratch + chain;
  ln.sd = kp.sd;
  ln.active = chain_raw < kp.C;
  const uint32_t tab = smem_u32(smem) + lane * 8;  // + row byte offset
  // [3][32] u32 value bits, then [3][32] u32 piece offsets
  unsigned long long* red_best = reinterpret_cast<unsigned long long*>(smem + kp.smem_red);
  unsigned long long* red_found = reinterpret_cast<unsigned long long*>(smem + kp.smem_red + 768);  // [2][32] u32 entry words (atomicMin), then [2][32] u32 piece info
  float* red_hyp = reinterpret_cast<float*>(smem + kp.smem_red + 1280);                       // [12][32]
  float* red_gc = reinterpret_cast<float*>(smem + kp.smem_red + 1280 + 12 * 128);             // [W][32]
  // the cluster's reductions live in rank 0's shared memory (shared::cluster addresses; rank 0 = this CTA when S == 1)
  const uint32_t a_best = cluster_map(smem_u32(red_best), 0), a_found = cluster_map(smem_u32(red_found), 0);
  const uint32_t a_cl = cluster_map(smem_u32(smem) + kp.smem_red_cl, 0);      // [kMaxSplit][32] f32
  const uint32_t a_partcl = cluster_map(smem_u32(smem) + kp.epi_cl, 0);       // [kMaxSplit][kPartRows][32] f32
  const uint32_t a_teamrows = cluster_map(smem_u32(smem) + kp.epi_team, 0);   // [T][2][32] f32
  constexpr uint32_t ESZ = CLIP ? (uint32_t)sizeof(EntryClip) : (uint32_t)sizeof(Entry);
  Ring ring;
  ring.init(smem_u32(smem) + kp.smem_ring + warp * (kStages * kp.stage_bytes),
            smem_u32(smem) + kp.smem_bar + warp * (kStages * 8), kp.stage_bytes, lane);
  const uint32_t b1_0 = wb.b1[vwarp], b1_1 = wb.b1[vwarp + 1];
  ring.begin(kp.stream1 + b1_0, b1_1 - b1_0);  // phase-1 pieces start streaming in while the prologue runs
  // Up to here the kernel has read only its parameters and the static plan: under programmatic dependent launch this
  // preamble (and the TMA of the first list stages) overlaps the tail of the kernel that produces theta.
  pdl_wait();
  pdl_launch_dependents();
  ln.th = after_wait(ln.th);
  ln.gr = after_wait(ln.gr);
  ln.sc = after_wait(ln.sc);
  // Everything the prologue needs from global memory is requested here, in one go, right after the ring has been
  // started (the ring's set-up ends in a fence, which would wait for loads already in flight): one round trip for all
  // of it.  (What the first pass over a team needs is then fetched one team ahead.)
  const int ndec = kp.ndec;
  const bool dc = ndec == 0;
  const float raw_in = ln.ld(o.raw);  // (its sigmoid -- a division, hence branches -- waits until the prologue's end)
  const bool lik = kp.lik_only != 0;
  const float u_in = (warp == 0 && !dc && !lik) ? ln.ld(o.u) : 0.0f;
  struct TeamIn { float za, zd, dz[4], xs[4]; int fl, v0, v1; };
  auto load_team = [&](int t) {
    TeamIn in;
    in.za = ln.ld(o.za + t);
    in.zd = ln.ld(o.zd + t);
#pragma unroll
    for (int i = 0; i < 4; i++) in.dz[i] = i < ndec ? ln.ld(o.dec[i] + t) : 0.0f;
#pragma unroll
    for (int j = 0; j < 4; j++) in.xs[j] = j < kp.K ? __ldg(kp.Xs + (size_t)t * kp.K + j) : 0.0f;
    in.fl = __ldg(kp.team_flags + t);
    in.v0 = __ldg(kp.team_vptr + t);
    in.v1 = __ldg(kp.team_vptr + t + 1);
    return in;
  };
  TeamIn nx = load_team(min(warp, kp.T - 1));
  float ba[4], bd[4];  // the first four covariate coefficients (more: the loop below)
#pragma unroll
  for (int j = 0; j < 4; j++) {
    ba[j] = j < kp.K ? ln.ld(o.beta_a + j) : 0.0f;
    bd[j] = j < kp.K ? ln.ld(o.beta_d + j) : 0.0f;
  }
  const Hyp hy_in = load_hyp_raw(kp, ln);
  {  // pull this CTA's slice of theta into L2 in one go: every later read of it is a hit
    const int nthr = W * 32;
    if (kp.sd == 1) {  // chain-major: 32 rows of D floats
      const int per = (kp.D + 31) / 32;
      for (int i = threadIdx.x; i < 32 * per; i += nthr) {
        const int c = min(group * kChains + i / per, kp.C - 1);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(kp.theta + (size_t)c * (size_t)kp.sc + (size_t)(i % per) * 32));
      }
    } else {
      for (int d = threadIdx.x; d < kp.D; d += nthr)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(kp.theta + (size_t)d * (size_t)kp.sd + (size_t)group * kChains));
    }
  }

  float lp_acc = 0.0f;  // this thread's share of the log-density
  float hacc = 0.0f;    // DIXON_COLES: d/d home_advantage

  // phase-2 slot sums per (side, team) when the plan deals a team's two sides to different warps (plan.cc)
  const bool p2s = kp.smem_p2 != 0 && S == 1;
  float* p2_tab = reinterpret_cast<float*>(smem + kp.smem_p2);
  uint32_t* red_clip = reinterpret_cast<uint32_t*>(smem + kp.smem_red_cl);  // [2][32] keys, this CTA's own (CLIP models)
  if (warp == 0) {
    const uint32_t zr = (uint32_t)kp.V * kRowBytes;  // zero rows used by padding entries
    if (kp.has1) {
      sts64(tab + kp.tabP1 + zr, 0.0f, 0.0f);
      sts64(tab + kp.tabQ1 + zr, 0.0f, 0.0f);
    }
    if (kp.has0) sts64(tab + kp.tabP0 + zr, 0.0f, 0.0f);
    if (CLIP) red_clip[lane] = red_clip[32 + lane] = 0u;
    if (crank == 0) {
      uint32_t* rb = reinterpret_cast<uint32_t*>(red_best);
      rb[lane] = rb[32 + lane] = rb[64 + lane] = 0u;                  // values
      rb[96 + lane] = rb[128 + lane] = rb[160 + lane] = 0xffffffffu;  // piece offsets (atomicMin)
      red_found[lane] = red_found[32 + lane] = ~0ull;
    }
  }
  if (CLIP) __syncthreads();  // the prologue's atomics on red_clip follow (all warps are still in step: cheap)

  BPLX_STAMP(0);
  // ---- prologue -------------------------------------------------------------------------------
  BPLX_STAMP_DEP(12, raw_in);
  {
    const Hyp hy = finish_hyp(kp, hy_in);
    BPLX_STAMP_DEP(13, hy.sig_a + hy.sig_d + hy.sig[0] + hy.sig[1] + hy.sig[2] + hy.sig[3] + nx.za + nx.zd);
    if (warp == 0) {  // the team pass reads them back from shared memory
      red_hyp[0 * 32 + lane] = hy.mu_d; red_hyp[1 * 32 + lane] = hy.sig_a; red_hyp[2 * 32 + lane] = hy.sig_d;
#pragma unroll
      for (int i = 0; i < 4; i++) { red_hyp[(3 + i) * 32 + lane] = hy.mu[i]; red_hyp[(7 + i) * 32 + lane] = hy.sig[i]; }
    }
    float exA = -FLT_MAX, exB = -FLT_MAX;  // largest exponent of either factor over this warp's virtual teams
    for (int t = warp; t < kp.T; t += W) {
      const TeamIn in = nx;
      if (t + W < kp.T) nx = load_team(t + W);
      if (p2s) {  // phase-2 slot tables of this team: a side without tau lists is never written
#pragma unroll
        for (int side = 0; side < 2; side++)
#pragma unroll
          for (int i = 0; i < 6; i++)
            if (i < 2 + ndec) p2_tab[((side * kp.T + t) * (2 + ndec) + i) * 32 + lane] = 0.0f;
      }
      float am = 0.0f, dm = hy.mu_d;
#pragma unroll
      for (int j = 0; j < 4; j++) {  // (absent covariates: x = 0, the sums do not change)
        am = fmaf(in.xs[j], ba[j], am);
        dm = fmaf(in.xs[j], bd[j], dm);
      }
      for (int k = 4; k < kp.K; k++) {
        const float x = __ldg(kp.Xs + (size_t)t * kp.K + k);
        am = fmaf(x, ln.ld(o.beta_a + k), am);
        dm = fmaf(x, ln.ld(o.beta_d + k), dm);
      }
      const float att = fmaf(in.za, hy.sig_a, am), def = fmaf(in.zd, hy.sig_d, dm);
      float x[4] = {dc ? hy.mu[0] : 0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (i < ndec) x[i] = fmaf(hy.sig[i], in.dz[i], hy.mu[i]);
      if (!(in.fl & 1) && ln.active && crank == 0) {  // team without matches: no list will write its slots
        *ln.g(o.za + t) = 0.0f;
        *ln.g(o.zd + t) = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (i < ndec) *ln.g(o.dec[i] + t) = 0.0f;
      }
      for (int v = in.v0; v < in.v1; v++) {
        const float cf = kp.Cf > 0 ? ln.ld(o.conf + __ldg(kp.v_conf + v)) : 0.0f;
        float ex[6];
        ex[eAh1] = att + x[0] + cf;
        ex[eBh1] = -def - x[2] - cf;
        ex[eBa1] = -def - x[3] - cf;
        ex[eAa1] = att + x[1] + cf;
        ex[eA0] = att + cf;
        ex[eB0] = -def - cf;
        const uint32_t r = (uint32_t)v * kRowBytes;
        if (kp.has1) {
          sts64(tab + kp.tabP1 + r, expf(ex[eAh1]), expf(ex[eBh1]));
          sts64(tab + kp.tabQ1 + r, expf(ex[eBa1]), expf(ex[eAa1]));
        }
        if (kp.has0) sts64(tab + kp.tabP0 + r, expf(ex[eA0]), expf(ex[eB0]));
        if (CLIP) {
          exA = fmaxf(exA, fmaxf(ex[eAh1], fmaxf(ex[eAa1], ex[eA0])));
          exB = fmaxf(exB, fmaxf(ex[eBh1], fmaxf(ex[eBa1], ex[eB0])));
        }
        if (!CLIP && crank == 0) {  // static sum of w * y * log(lambda): linear in the exponents
#pragma unroll
          for (int e = 0; e < 6; e++) lp_acc = fmaf(__ldg(kp.yexp + (size_t)v * 6 + e), ex[e], lp_acc);
        }
      }
    }
    if (warp == 0) red_hyp[11 * 32 + lane] = (dc || lik) ? 0.5f : sigmoid_clipped(u_in);
    if (CLIP && warp < kp.T) {
      atomicMax(red_clip + lane, float_key(exA));
      atomicMax(red_clip + 32 + lane, float_key(exB));
    }
  }
  const float r = lik ? raw_in : sigmoid_clipped(raw_in);  // (likelihood-only: corr_coef_raw itself comes in)
  BPLX_STAMP(1);
  sync_all();

  BPLX_STAMP(2);
  // every rate is exp(A-side exponent of one virtual team + B-side exponent of another): with the two maxima below
  // log 15 (less a margin for the rounding of exp and the product) no rate of the chain can be at the clip
  bool fast1 = false;
  if (CLIP) fast1 = __all_sync(kFull, key_float(red_clip[lane]) + key_float(red_clip[32 + lane]) < 2.707f) && !(kp.force_clip_forms & 1);
  // ---- phase 1 ----------------------------------------------------------------------------------
  float best[3] = {0.0f, 0.0f, 0.0f};
  uint32_t besth[3] = {0u, 0u, 0u};  // byte offset (in stream1) of the header of the piece that holds the maximum
  {
    float g[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f}, cacc = 0.0f;
    const uint32_t nst = ring.num_stages();
    for (uint32_t k = 0; k < nst; k++) {
      uint32_t bytes;
      const uint32_t a0 = ring.acquire(k, &bytes);
      if (k == 0) BPLX_STAMP(16);
      uint32_t a = a0;
      const uint32_t aend = a0 + bytes;
      while (a + kp.min_piece1 <= aend) {
        const uint32_t hoff = b1_0 + k * ring.S + (a - a0);
        const Hdr L = unpack_hdr(lds128u(a));
        a += 16;
        const uint32_t e_end = a + L.n0 * ESZ;
        if (L.flags & kTeamFirst) {
#pragma unroll
          for (int e = 0; e < 6; e++) g[e] = 0.0f;
        }
        float2 own = lds64(tab + L.own_off);
        if (L.kind >= kH0) { const float s = own.x; own.x = own.y; own.y = s; }
        const bool home = (L.kind & 1) == 0;
        float gx, gy;
        if (!CLIP) {
          float2 acc0 = make_float2(0.0f, 0.0f), acc1 = acc0;  // sum w * (X-side row, Y-side row), even / odd entries
          if (home) {
            float m1 = 0.0f, m2 = 0.0f, m3 = 0.0f;
            walk16(a, e_end, [&](const uint32_t at) {
              const uint4 q = lds128u(at);  // two entries
              const float2 ea = lds64(tab + q.x), eb = lds64(tab + q.z);
              acc0 = fma2(bc2(__uint_as_float(q.y)), ea, acc0);
              acc1 = fma2(bc2(__uint_as_float(q.w)), eb, acc1);
              m1 = fmaxf(m1, fmaxf(ea.x, eb.x));
              m2 = fmaxf(m2, fmaxf(ea.y, eb.y));
              m3 = fmaxf(m3, fmaxf(ea.x * ea.y, eb.x * eb.y));
            });
            const float v0 = own.x * m1, v1 = own.y * m2, v2 = (own.x * own.y) * m3;
            if (v0 > best[0]) { best[0] = v0; besth[0] = hoff; }
            if (v1 > best[1]) { best[1] = v1; besth[1] = hoff; }
            if (v2 > best[2]) { best[2] = v2; besth[2] = hoff; }
          } else {
            walk16(a, e_end, [&](const uint32_t at) {
              const uint4 q = lds128u(at);
              const float2 ea = lds64(tab + q.x), eb = lds64(tab + q.z);
              acc0 = fma2(bc2(__uint_as_float(q.y)), ea, acc0);
              acc1 = fma2(bc2(__uint_as_float(q.w)), eb, acc1);
            });
          }
          const float SX = own.x * (acc0.x + acc1.x), SY = own.y * (acc0.y + acc1.y);
          lp_acc -= 0.5f * (SX + SY);  // every match is in two lists
          gx = -SX;
          gy = -SY;
        } else {
          float lp = 0.0f, m1 = 0.0f, m2 = 0.0f, m3 = 0.0f;
          if (fast1) {
            if (home) rate_piece_clip<true, true>(a, e_end, own, tab, lp, m1, m2, m3, gx, gy);
            else rate_piece_clip<true, false>(a, e_end, own, tab, lp, m1, m2, m3, gx, gy);
          } else {
            if (home) rate_piece_clip<false, true>(a, e_end, own, tab, lp, m1, m2, m3, gx, gy);
            else rate_piece_clip<false, false>(a, e_end, own, tab, lp, m1, m2, m3, gx, gy);
          }
          if (home) {
            lp_acc += lp;
            if (m1 > best[0]) { best[0] = m1; besth[0] = hoff; }
            if (m2 > best[1]) { best[1] = m2; besth[1] = hoff; }
            if (m3 > best[2]) { best[2] = m3; besth[2] = hoff; }
          }
        }
        add_own(g, L.kind, gx, gy);
        if (kp.Cf > 0) {
          cacc += L.kind == kH1 ? gx - gy : gy - gx;  // d/d (A - B) of the virtual team
          if (L.flags & kVteamLast) {
            if (ln.active) ln.sc[(size_t)L.vteam * kp.Cpad] = cacc;
            cacc = 0.0f;
          }
        }
        if (L.flags & kTeamLast) put_raw<false>(kp, ln, (int)L.team, g, hacc);
        if (k == 0 && a - a0 <= kp.stage_bytes / 2) BPLX_STAMP(17);
      }
      ring.release(k);
    }
  }
  const uint32_t b2_0 = wb.b2[vwarp], b2_1 = wb.b2[vwarp + 1];
  ring.begin(kp.stream2 + b2_0, b2_1 - b2_0);  // tau pieces start streaming in during the bounds step

  BPLX_STAMP(3);
  // ---- bounds (bpl/_util.py:17-31) ------------------------------------------------------------------
  // the maxima and the piece each came from, with 32-bit atomics in two steps: the value first, then the owners of the
  // maximum agree on the lowest piece offset.  (A 64-bit max on value | offset is a CAS loop in shared memory -- about
  // 1 us with 20 warps on the same 32 words -- and in a cluster the local CAS loop and the remote atomic do not
  // exclude each other: measured lost updates.)
#pragma unroll
  for (int q = 0; q < 3; q++)
    if (best[q] > 0.0f) dsm_atom_max_u32(a_best + (uint32_t)(q * 32 + lane) * 4u, __float_as_uint(best[q]));
  sync_all();
  BPLX_STAMP(14);
#pragma unroll
  for (int q = 0; q < 3; q++) {
    const float gq = __uint_as_float(dsm_ld_u32(a_best + (uint32_t)(q * 32 + lane) * 4u));
    if (best[q] == gq && gq > 0.0f) dsm_atom_min_u32(a_best + (uint32_t)((3 + q) * 32 + lane) * 4u, besth[q]);
    best[q] = gq;
  }
  sync_all();
  BPLX_STAMP(15);
#pragma unroll
  for (int q = 0; q < 3; q++) besth[q] = dsm_ld_u32(a_best + (uint32_t)((3 + q) * 32 + lane) * 4u);
  // no rate of these 32 chains at the clip: phase 2 takes the short form of the clipped arithmetic
  const bool noclip = CLIP && __all_sync(kFull, best[0] < 15.0f && best[1] < 15.0f) && !(kp.force_clip_forms & 2);
  const float Lam = fmaxf(best[0], best[1]);
  const int qlam = best[0] >= best[1] ? 0 : 1;
  // arg-max search, loads first: the headers of each chain's two arg-max pieces and this warp's first candidate entry
  // of each (its address needs only the piece offset; past the end of a short piece it reads a neighbour or the zero
  // slack behind the stream, and is ignored) -- one round trip to L2 for all four
  bool s_need[2];
  uint32_t s_hoff[2], s_off0[2];
  uint4 s_hdr[2];
#pragma unroll
  for (int which = 0; which < 2; which++) {
    s_need[which] = (which == 0 || best[2] > 1.0f) && best[which == 0 ? qlam : 2] > 0.0f;  // UB = 1: no dependence on the rates
    s_hoff[which] = s_need[which] ? (which == 0 ? (qlam == 0 ? besth[0] : besth[1]) : besth[2]) : b1_0;
    const unsigned char* base = kp.stream1 + s_hoff[which];
    s_off0[which] = __ldg(reinterpret_cast<const uint32_t*>(base + 16 + (size_t)vwarp * ESZ));
    s_hdr[which] = __ldg(reinterpret_cast<const uint4*>(base));
  }
  const float LB = -1.0f / Lam;
  const float UB = fminf(1.0f / best[2], 1.0f);
  const float cc = fmaf(r, UB - LB, LB);

  BPLX_STAMP(4);
  // ---- arg-max search: virtual warp w looks at entries w, w+VW, ... of each chain's two arg-max pieces -----------
#pragma unroll
  for (int which = 0; which < 2; which++) {
    const bool need = s_need[which];
    const float target = which == 0 ? Lam : best[2];
    const int q = which == 0 ? qlam : 2;
    const unsigned char* ent = kp.stream1 + s_hoff[which] + 16;
    uint32_t off_next = s_off0[which];
    const Hdr L = unpack_hdr(s_hdr[which]);
    const uint32_t n = need ? L.n0 : 0u;
    float2 own = lds64(tab + (need ? L.own_off : 0u));
    if (L.kind >= kH0) { const float s = own.x; own.x = own.y; own.y = s; }
    const uint32_t nmax = __reduce_max_sync(kFull, n);
    uint32_t found = 0xffffffffu, info = 0u;
#pragma unroll 1
    for (uint32_t i = vwarp; i < nmax; i += VW) {
      const uint32_t off = off_next;
      if (i + VW < nmax) off_next = __ldg(reinterpret_cast<const uint32_t*>(ent + (size_t)(i + VW) * ESZ));
      if (i < n) {
        const float2 ea = lds64(tab + off);
        const float X = own.x * ea.x, Y = own.y * ea.y;
        float val;
        if (CLIP) {
          const float Xc = fminf(X, 15.0f), Yc = fminf(Y, 15.0f);
          val = q == 0 ? Xc : (q == 1 ? Yc : Xc * Yc);
        } else {
          val = q == 0 ? X : (q == 1 ? Y : (own.x * own.y) * (ea.x * ea.y));
        }
        if (val == target && found == 0xffffffffu) {
          // entry index | "X not clipped" | "Y not clipped" | opponent row offset (< 4 MB)
          found = (i << 24) | ((!CLIP || X < 15.0f) ? 1u << 23 : 0u) | ((!CLIP || Y < 15.0f) ? 1u << 22 : 0u) | off;
          info = L.vteam | (L.kind << 16);  // own vteam | kind: the same for every finder of this chain
        }
      }
    }
    // several warps find an entry only under exact ties inside the piece (clipped rates, or rates that overflowed to
    // inf far from the typical set): the lowest entry index wins -- atomicMin, so the choice does not depend on timing
    if (found != 0xffffffffu) {
      dsm_atom_min_u32(a_found + (uint32_t)(which * 32 + lane) * 4u, found);
      dsm_st_u32(a_found + (uint32_t)(64 + which * 32 + lane) * 4u, info);
    }
  }

  BPLX_STAMP(5);
  // ---- phase 2: tau terms (bpl/_util.py:54-91) -----------------------------------------------------
  float gc = 0.0f;
  {
    float g[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f}, cacc = 0.0f;
    const uint32_t nst = ring.num_stages();
    for (uint32_t k = 0; k < nst; k++) {
      uint32_t bytes;
      const uint32_t a0 = ring.acquire(k, &bytes);
      if (k == 0) BPLX_STAMP(18);
      uint32_t a = a0;
      const uint32_t aend = a0 + bytes;
      while (a + kp.min_piece2 <= aend) {
        const Hdr L = unpack_hdr(lds128u(a));
        a += 16;
        if (L.flags & kTeamFirst) {
#pragma unroll
          for (int e = 0; e < 6; e++) g[e] = 0.0f;
        }
        float2 own = lds64(tab + L.own_off);
        if (L.kind >= kH0) { const float s = own.x; own.x = own.y; own.y = s; }
        const bool home = (L.kind & 1) == 0;
        float lt, du, gx, gy;
        if (!CLIP) tau_piece<kTauPlain>(a, L, own, home, cc, tab, lt, du, gx, gy);
        else if (noclip) tau_piece<kTauUnclipped>(a, L, own, home, cc, tab, lt, du, gx, gy);
        else tau_piece<kTauClipped>(a, L, own, home, cc, tab, lt, du, gx, gy);
        if (home) {
          lp_acc = fmaf(lt, kLn2, lp_acc);
          gc += du;
        }
        add_own(g, L.kind, gx, gy);
        if (kp.Cf > 0) {
          cacc += L.kind == kH1 ? gx - gy : gy - gx;
          if (L.flags & kVteamLast) {
            if (ln.active) red_add(ln.sc + (size_t)L.vteam * kp.Cpad, cacc);
            cacc = 0.0f;
          }
        }
        if (L.flags & kTeamLast) {
          if (p2s) put_side(kp, p2_tab + (((L.kind & 1) * kp.T + L.team) * (2 + ndec)) * 32 + lane, g, hacc);
          else put_raw<true>(kp, ln, (int)L.team, g, hacc);
        }
      }
      ring.release(k);
    }
  }
  BPLX_STAMP(6);
  red_gc[warp * 32 + lane] = gc;
  __syncthreads();
  gc = 0.0f;
  for (int w = 0; w < W; w++) gc += red_gc[w * 32 + lane];
  if (S > 1) {  // the CTAs' sums meet in rank 0
    if (warp == 0) dsm_st_f32(a_cl + (uint32_t)(crank * 32 + lane) * 4u, gc);
    cg::this_cluster().sync();
    gc = 0.0f;
    for (int q = 0; q < S; q++) gc += dsm_ld_f32(a_cl + (uint32_t)(q * 32 + lane) * 4u);
  }
  // tables are dead from here on; raw slots hold both phases
  {  // the 1-1 matches: tau = 1 - c for all of them
    const float t11 = fmaxf(1.0f - cc, 0.0f);
    gc -= kp.w11 / t11;
    if (vwarp == 0 && kp.w11 != 0.0f) lp_acc = fmaf(kp.w11, logf(t11), lp_acc);
  }

  BPLX_STAMP(7);
  // ---- arg-max fix-up: every warp works out the two matches of its chain and folds
  //      them into the team pass below ---------------------------------------------------------------------
  Fixup fx;
  fx.h1 = 0u;
  fx.confs = 0u;
#pragma unroll
  for (int which = 0; which < 2; which++) {
    const uint32_t packed = dsm_ld_u32(a_found + (uint32_t)(which * 32 + lane) * 4u);
    fx.teams[which] = 0xffffffffu;  // (unused here: the team pass matches virtual-team ranges, no v_team look-up)
    fx.vts[which] = 0xffffffffu;    // none
    fx.vx[which] = fx.vy[which] = 0.0f;
    if (packed != 0xffffffffu) {  // else: UB = 1 (or nothing matched: cannot happen, same arithmetic as phase 1)
      const uint32_t info = dsm_ld_u32(a_found + (uint32_t)(64 + which * 32 + lane) * 4u);
      const uint32_t f_off = packed & 0x3fffffu;
      const bool h1 = ((info >> 16) & 3u) == kH1;
      const uint32_t own_v = info & 0xffffu;
      const uint32_t opp_v = (f_off - (h1 ? kp.tabQ1 : kp.tabP0)) / kRowBytes;
      const bool xfree = (packed >> 23) & 1u, yfree = (packed >> 22) & 1u;
      if (which == 0) {
        const float wgt = gc * (1.0f - r) / Lam;  // dc/dLB * dLB/d eta
        fx.vx[0] = (qlam == 0 && xfree) ? wgt : 0.0f;
        fx.vy[0] = (qlam == 1 && yfree) ? wgt : 0.0f;
      } else {
        const float wgt = -gc * r / best[2];  // dc/dUB * dUB/d eta (UB = 1 / max lambda_h lambda_a)
        fx.vx[1] = xfree ? wgt : 0.0f;
        fx.vy[1] = yfree ? wgt : 0.0f;
      }
      fx.h1 |= (h1 ? 1u : 0u) << which;
      fx.vts[which] = own_v | (opp_v << 16);
      if (kp.Cf > 0)
        fx.confs |= ((uint32_t)__ldg(kp.v_conf + own_v) | ((uint32_t)__ldg(kp.v_conf + opp_v) << 8)) << (16 * which);
    }
  }

  BPLX_STAMP(8);
  // ---- team pass: raw slots -> parameter gradients, priors, hyper-parameter sums ---------------------------
  const bool has_rho = !dc && !lik;
  const float pri = lik ? 0.0f : 1.0f;  // likelihood-only: no prior terms
  float u = 0.5f, rho = 0.0f, inv_s2 = 1.0f;
  if (has_rho) {
    u = red_hyp[11 * 32 + lane];
    rho = 2.0f * u - 1.0f;
    inv_s2 = 1.0f / (1.0f - rho * rho);
  }
  float a_mu_d = 0.0f, a_ls_a = 0.0f, a_ls_d = 0.0f, a_rho = 0.0f, a_mu[4], a_ls[4];
#pragma unroll
  for (int i = 0; i < 4; i++) a_mu[i] = a_ls[i] = 0.0f;
  {
    Hyp hy;
    hy.mu_d = red_hyp[0 * 32 + lane]; hy.sig_a = red_hyp[1 * 32 + lane]; hy.sig_d = red_hyp[2 * 32 + lane];
#pragma unroll
    for (int i = 0; i < 4; i++) { hy.mu[i] = red_hyp[(3 + i) * 32 + lane]; hy.sig[i] = red_hyp[(7 + i) * 32 + lane]; }

    for (int t = vwarp; t < kp.T; t += VW) {
      const float za = ln.ld(o.za + t), zd = ln.ld(o.zd + t);
      const uint32_t v0 = (uint32_t)__ldg(kp.team_vptr + t), nv = (uint32_t)__ldg(kp.team_vptr + t + 1) - v0;
      const float4 ys = __ldg(reinterpret_cast<const float4*>(kp.yteam + (size_t)t * 8));
      const float2 ys2 = __ldg(reinterpret_cast<const float2*>(kp.yteam + (size_t)t * 8 + 4));
      float ra = ld_cg(ln.g(o.za + t)) + ys.x;
      float rd = ld_cg(ln.g(o.zd + t)) + ys.y;
      float rx[4] = {ys.z, ys.w, ys2.x, ys2.y};
      float dec[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (i < ndec) {
          dec[i] = ln.ld(o.dec[i] + t);
          rx[i] += ld_cg(ln.g(o.dec[i] + t));
        }
      }
      if (p2s) {  // home-side sums, then away-side sums
#pragma unroll
        for (int side = 0; side < 2; side++) {
          const float* q = p2_tab + ((side * kp.T + t) * (2 + ndec)) * 32 + lane;
          ra += q[0];
          rd += q[32];
#pragma unroll
          for (int i = 0; i < 4; i++)
            if (i < ndec) rx[i] += q[(2 + i) * 32];
        }
      }
#pragma unroll
      for (int which = 0; which < 2; which++) {
        // the arg-max match's own / opponent virtual team is one of this team's
        if ((fx.vts[which] & 0xffffu) - v0 < nv) fold_fixup(fx, which, 0, ra, rd, rx);
        if ((fx.vts[which] >> 16) - v0 < nv) fold_fixup(fx, which, 1, ra, rd, rx);
      }
      float p_za, p_zd;
      if (has_rho) {  // za ~ N(0,1), zd ~ N(rho za, sqrt(1-rho^2))  (extended_dixon_coles.py:165-174)
        const float e = zd - rho * za;
        const float es = e * inv_s2;
        lp_acc -= 0.5f * (za * za + e * es);
        p_za = -za + rho * es;
        p_zd = -es;
        a_rho += es * za - rho * es * es + rho * inv_s2;
      } else {
        lp_acc -= pri * 0.5f * (za * za + zd * zd);
        p_za = -pri * za;
        p_zd = -pri * zd;
      }
      if (ln.active) {
        *ln.g(o.za + t) = fmaf(hy.sig_a, ra, p_za);
        *ln.g(o.zd + t) = fmaf(hy.sig_d, rd, p_zd);
      }
      a_ls_a = fmaf(hy.sig_a * za, ra, a_ls_a);
      a_ls_d = fmaf(hy.sig_d * zd, rd, a_ls_d);
      a_mu_d += rd;
      if (dc) a_mu[0] += rx[0];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (i < ndec) {
          lp_acc -= pri * 0.5f * dec[i] * dec[i];
          if (ln.active) *ln.g(o.dec[i] + t) = fmaf(hy.sig[i], rx[i], -pri * dec[i]);
          a_mu[i] += rx[i];
          a_ls[i] = fmaf(hy.sig[i] * dec[i], rx[i], a_ls[i]);
        }
      }
      if (kp.K > 0) {  // rows for the covariate-coefficient pass
        dsm_st_f32(a_teamrows + (uint32_t)(t * 64 + lane) * 4u, ra);  // (rank 0's rows)
        dsm_st_f32(a_teamrows + (uint32_t)(t * 64 + 32 + lane) * 4u, rd);
      }
    }
    if (dc) a_mu[0] += hacc;
  }
  // confederation strengths: N(0,1) prior + sum over the virtual teams of the confederation
  for (int k = vwarp; k < kp.Cf; k += VW) {
    const float cf = ln.ld(o.conf + k);
    float s = __ldg(kp.yconf + k) - pri * cf;
    lp_acc -= pri * 0.5f * cf * cf;
    const int j0 = __ldg(kp.conf_vptr + k), j1 = __ldg(kp.conf_vptr + k + 1);
    for (int j = j0; j < j1; j++) s += ld_cg(ln.sc + (size_t)__ldg(kp.conf_vlist + j) * kp.Cpad);
#pragma unroll
    for (int ws = 0; ws < 4; ws++)
      if (fx.vts[ws >> 1] != 0xffffffffu && ((fx.confs >> (8 * ((ws >> 1) * 2 + (ws & 1)))) & 0xffu) == (uint32_t)k)
        s += fixup_conf(fx, ws >> 1, ws & 1);
    if (ln.active) *ln.g(o.conf + k) = s;
  }

  BPLX_STAMP(9);
  // ---- cross-warp reduction of the hyper accumulators (table area is reused) ----------------------------
  float* part = reinterpret_cast<float*>(smem + kp.epi_part);
  {
    float* p = part + (size_t)warp * kPartRows * 32 + lane;
    p[0 * 32] = lp_acc; p[1 * 32] = a_mu_d; p[2 * 32] = a_ls_a; p[3 * 32] = a_ls_d;
#pragma unroll
    for (int i = 0; i < 4; i++) { p[(4 + i) * 32] = a_mu[i]; p[(8 + i) * 32] = a_ls[i]; }
    p[12 * 32] = a_rho;
  }
  __syncthreads();
  if (S > 1) {  // per-CTA sums of the 13 rows go to rank 0; the other CTAs are done
    for (int row = warp; row < kAccRows; row += W) {
      float s = 0.0f;
      for (int w = 0; w < W; w++) s += part[((size_t)w * kPartRows + row) * 32 + lane];
      dsm_st_f32(a_partcl + (uint32_t)((crank * kPartRows + row) * 32 + lane) * 4u, s);
    }
    cg::this_cluster().sync();
    if (crank != 0) return;
  }
  BPLX_STAMP(10);
  // the epilogue items are dealt round-robin to the warps: scalar hyper sites, u, corr_coef_raw, coefficients
  auto total = [&](int row) {
    float s = 0.0f;
    if (S > 1) {
      for (int q = 0; q < S; q++) s += dsm_ld_f32(a_partcl + (uint32_t)((q * kPartRows + row) * 32 + lane) * 4u);
    } else {
      for (int w = 0; w < W; w++) s += part[((size_t)w * kPartRows + row) * 32 + lane];
    }
    return s;
  };
  float lp = 0.0f;
  const int n_items = kp.nhyper + 2 + 2 * kp.K;
  for (int item = warp; item < n_items; item += W) {
    if (item < kp.nhyper) {
      const HyperDesc hd = kp.hyper[item];
      const float x = ln.ld(hd.off);
      const float acc = total(hd.row);
      float gval;
      if (hd.kind == 0) {  // Normal(loc, scale)
        const float z = (x - hd.loc) * hd.inv_scale;
        lp -= 0.5f * z * z;
        gval = fmaf(-z, hd.inv_scale, acc);
      } else if (hd.kind == 2) {  // likelihood-only: no prior
        gval = acc;
      } else {  // HalfNormal(scale) on exp(x) + Jacobian x
        const float z = expf(x) * hd.inv_scale;
        lp += fmaf(-0.5f * z, z, x);
        gval = fmaf(-z, z, 1.0f) + acc;
      }
      if (ln.active) *ln.g(hd.off) = gval;
    } else if (item == kp.nhyper) {
      if (has_rho) {  // u ~ Beta(2,4) + sigmoid Jacobian; rho = 2u - 1; sum_t -log sqrt(1 - rho^2)
        lp += 0.5f * (float)kp.T * logf(inv_s2) + 2.0f * logf(u) + 4.0f * logf(1.0f - u);
        if (ln.active) *ln.g(o.u) = 2.0f - 6.0f * u + total(12) * 2.0f * u * (1.0f - u);
      }
    } else if (item == kp.nhyper + 1) {
      // corr_coef_raw ~ Beta(2,2) + Jacobian; corr_coef = LB + r (UB - LB); also the likelihood + team-prior sum
      lp += total(0) + (lik ? 0.0f : 2.0f * (logf(r) + logf(1.0f - r)));
      if (ln.active) {
        *ln.g(o.raw) = lik ? gc * (UB - LB) : 2.0f * (1.0f - 2.0f * r) + gc * r * (1.0f - r) * (UB - LB);
        if (kp.corr_coef) kp.corr_coef[chain] = cc;
      }
    } else {  // covariate coefficients: d/d beta[k] = sum_t Xs[t,k] * d/d (att | def)[t]; N(0,1) prior
      const int task = item - kp.nhyper - 2, k = task >> 1, isd = task & 1;
      const float* rows = reinterpret_cast<const float*>(smem + kp.epi_team) + isd * 32 + lane;
      const int d = (isd ? o.beta_d : o.beta_a) + k;
      const float b = ln.ld(d);
      float s = -b;
      lp -= 0.5f * b * b;
      for (int t = 0; t < kp.T; t++) s = fmaf(__ldg(kp.Xs + (size_t)t * kp.K + k), rows[(size_t)t * 64], s);
      if (ln.active) *ln.g(d) = s;
    }
  }
  BPLX_STAMP(11);
  red_gc[warp * 32 + lane] = lp;  // every warp read its gc sum before the previous barrier: the rows are free
  __syncthreads();
  if (warp == 0 && ln.active) {
    lp = kp.const_term;
    for (int w = 0; w < W; w++) lp += red_gc[w * 32 + lane];
    // a log-density is never +inf: that is an intermediate that overflowed float32 far from the typical set (a sampler
    // would accept such a point as the best ever seen); NaN is what the callers reject
    kp.lp[chain] = lp == INFINITY ? NAN : lp;
  }
}

// ---- host launcher --------------------------------------------------------------------------------------
template <bool CLIP>
static int launch_t(const KernelParams& kp, const WarpBounds& wb, cudaStream_t stream, bool set_attr, bool pdl) {
  auto* fn = &logdensity_kernel<CLIP>;
  if (set_attr) {
    BPLX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return BPLX_OK;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(kp.ngroups * kp.split));
  cfg.blockDim = dim3((unsigned)(kp.nwarps * 32));
  cfg.dynamicSmemBytes = kp.smem_total;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (kp.split > 1) {  // one thread-block cluster per group of chains
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = (unsigned)kp.split;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    na++;
  }
  if (pdl && pdl_enabled()) {  // the kernel's preamble may overlap the tail of its predecessor in the stream (pdl_wait() inside)
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    na++;
  }
  cfg.attrs = at;
  cfg.numAttrs = (unsigned)na;
  BPLX_CUDA(cudaLaunchKernelEx(&cfg, fn, kp, wb));
  BPLX_CUDA(cudaGetLastError());
  note_launch(1);
  return BPLX_OK;
}

template <bool CLIP>
static int max_clusters_t(const KernelParams& kp, int split) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)split);
  cfg.blockDim = dim3((unsigned)(kp.nwarps * 32));
  cfg.dynamicSmemBytes = kp.smem_total;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)split;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, &logdensity_kernel<CLIP>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int logdensity_max_clusters(const KernelParams& kp, int split) {
  return kp.clip ? max_clusters_t<true>(kp, split) : max_clusters_t<false>(kp, split);
}

#ifdef BPLX_TIMELINE
extern "C" int bplx_timeline_read(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_timeline, sizeof(g_timeline)) == cudaSuccess ? 0 : -1;
}
#endif

int logdensity_set_attributes(const KernelParams& kp) {
  static const WarpBounds none{};
  return kp.clip ? launch_t<true>(kp, none, nullptr, true, false) : launch_t<false>(kp, none, nullptr, true, false);
}
int launch_logdensity(const KernelParams& kp, const WarpBounds& wb, cudaStream_t stream, bool pdl) {
  return kp.clip ? launch_t<true>(kp, wb, stream, false, pdl) : launch_t<false>(kp, wb, stream, false, pdl);
}

}  // namespace bplx
