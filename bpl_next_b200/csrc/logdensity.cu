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

#include "k1_common.cuh"
#include "problem.h"

namespace cg = cooperative_groups;

namespace bplx {

template <bool CLIP>
__global__ void __launch_bounds__(kMaxWarps * 32, 1) logdensity_kernel(const __grid_constant__ KernelParams kp) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = kp.nwarps;
  const ThetaOffsets& o = kp.off;
  const int S = kp.split;  // CTAs of the cluster that shares this group of chains (1: no cluster)
  const int crank = S > 1 ? (int)cg::this_cluster().block_rank() : 0;
  const int group = (int)blockIdx.x / S;
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
  // one CTA: [3][32] u64 (value bits << 32 | piece offset).  Cluster: [3][32] u32 value bits, then [3][32] u32 piece offsets
  unsigned long long* red_best = reinterpret_cast<unsigned long long*>(smem + kp.smem_red);
  unsigned long long* red_found = reinterpret_cast<unsigned long long*>(smem + kp.smem_red + 768);  // [2][32] entry | info << 32
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
  const uint32_t b1_0 = __ldg(kp.warp_b1 + vwarp), b1_1 = __ldg(kp.warp_b1 + vwarp + 1);
  ring.begin(kp.stream1 + b1_0, b1_1 - b1_0);  // phase-1 pieces start streaming in while the prologue runs
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

  const int ndec = kp.ndec;
  const bool dc = ndec == 0;
  float lp_acc = 0.0f;  // this thread's share of the log-density
  float hacc = 0.0f;    // DIXON_COLES: d/d home_advantage

  // ---- prologue -------------------------------------------------------------------------------
  if (warp == 0) {
    const uint32_t zr = (uint32_t)kp.V * kRowBytes;  // zero rows used by padding entries
    if (kp.has1) {
      sts64(tab + kp.tabP1 + zr, 0.0f, 0.0f);
      sts64(tab + kp.tabQ1 + zr, 0.0f, 0.0f);
    }
    if (kp.has0) sts64(tab + kp.tabP0 + zr, 0.0f, 0.0f);
    if (crank == 0) {
      if (S == 1) {
        red_best[lane] = red_best[32 + lane] = red_best[64 + lane] = 0ull;
      } else {
        uint32_t* rb = reinterpret_cast<uint32_t*>(red_best);
        rb[lane] = rb[32 + lane] = rb[64 + lane] = 0u;                         // values
        rb[96 + lane] = rb[128 + lane] = rb[160 + lane] = 0xffffffffu;         // piece offsets (atomicMin)
      }
      red_found[lane] = red_found[32 + lane] = ~0ull;
    }
  }
  const float r = sigmoid_clipped(ln.ld(o.raw));
  {
    const Hyp hy = load_hyp(kp, ln);
    if (warp == 0) {  // the team pass reads them back from shared memory
      red_hyp[0 * 32 + lane] = hy.mu_d; red_hyp[1 * 32 + lane] = hy.sig_a; red_hyp[2 * 32 + lane] = hy.sig_d;
#pragma unroll
      for (int i = 0; i < 4; i++) { red_hyp[(3 + i) * 32 + lane] = hy.mu[i]; red_hyp[(7 + i) * 32 + lane] = hy.sig[i]; }
      red_hyp[11 * 32 + lane] = dc ? 0.5f : sigmoid_clipped(ln.ld(o.u));
    }
    for (int t = warp; t < kp.T; t += W) {
      float am = 0.0f, dm = hy.mu_d;
      for (int k = 0; k < kp.K; k++) {
        const float x = __ldg(kp.Xs + (size_t)t * kp.K + k);
        am = fmaf(x, ln.ld(o.beta_a + k), am);
        dm = fmaf(x, ln.ld(o.beta_d + k), dm);
      }
      const float att = fmaf(ln.ld(o.za + t), hy.sig_a, am), def = fmaf(ln.ld(o.zd + t), hy.sig_d, dm);
      float x[4] = {dc ? hy.mu[0] : 0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (i < ndec) x[i] = fmaf(hy.sig[i], ln.ld(o.dec[i] + t), hy.mu[i]);
      if (!(__ldg(kp.team_flags + t) & 1) && ln.active && crank == 0) {  // team without matches: no list will write its slots
        *ln.g(o.za + t) = 0.0f;
        *ln.g(o.zd + t) = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (i < ndec) *ln.g(o.dec[i] + t) = 0.0f;
      }
      const int v0 = __ldg(kp.team_vptr + t), v1 = __ldg(kp.team_vptr + t + 1);
      for (int v = v0; v < v1; v++) {
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
        if (!CLIP && crank == 0) {  // static sum of w * y * log(lambda): linear in the exponents
#pragma unroll
          for (int e = 0; e < 6; e++) lp_acc = fmaf(__ldg(kp.yexp + (size_t)v * 6 + e), ex[e], lp_acc);
        }
      }
    }
  }
  sync_all();

  // ---- phase 1 ----------------------------------------------------------------------------------
  float best[3] = {0.0f, 0.0f, 0.0f};
  uint32_t besth[3] = {0u, 0u, 0u};  // byte offset (in stream1) of the header of the piece that holds the maximum
  {
    float g[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f}, cacc = 0.0f;
    const uint32_t nst = ring.num_stages();
    for (uint32_t k = 0; k < nst; k++) {
      uint32_t bytes;
      const uint32_t a0 = ring.acquire(k, &bytes);
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
          float ax0 = 0.0f, ay0 = 0.0f, ax1 = 0.0f, ay1 = 0.0f;
          if (home) {
            float m1 = 0.0f, m2 = 0.0f, m3 = 0.0f;
#pragma unroll 4
            for (; a < e_end; a += 16) {
              const uint4 q = lds128u(a);  // two entries
              const float2 ea = lds64(tab + q.x), eb = lds64(tab + q.z);
              const float wa = __uint_as_float(q.y), wb = __uint_as_float(q.w);
              ax0 = fmaf(wa, ea.x, ax0); ay0 = fmaf(wa, ea.y, ay0);
              ax1 = fmaf(wb, eb.x, ax1); ay1 = fmaf(wb, eb.y, ay1);
              m1 = fmaxf(m1, fmaxf(ea.x, eb.x));
              m2 = fmaxf(m2, fmaxf(ea.y, eb.y));
              m3 = fmaxf(m3, fmaxf(ea.x * ea.y, eb.x * eb.y));
            }
            const float v0 = own.x * m1, v1 = own.y * m2, v2 = (own.x * own.y) * m3;
            if (v0 > best[0]) { best[0] = v0; besth[0] = hoff; }
            if (v1 > best[1]) { best[1] = v1; besth[1] = hoff; }
            if (v2 > best[2]) { best[2] = v2; besth[2] = hoff; }
          } else {
#pragma unroll 4
            for (; a < e_end; a += 16) {
              const uint4 q = lds128u(a);
              const float2 ea = lds64(tab + q.x), eb = lds64(tab + q.z);
              const float wa = __uint_as_float(q.y), wb = __uint_as_float(q.w);
              ax0 = fmaf(wa, ea.x, ax0); ay0 = fmaf(wa, ea.y, ay0);
              ax1 = fmaf(wb, eb.x, ax1); ay1 = fmaf(wb, eb.y, ay1);
            }
          }
          const float SX = own.x * (ax0 + ax1), SY = own.y * (ay0 + ay1);
          lp_acc -= 0.5f * (SX + SY);  // every match is in two lists
          gx = -SX;
          gy = -SY;
        } else {
          gx = gy = 0.0f;
          float lp2 = 0.0f, lpw = 0.0f, m1 = 0.0f, m2 = 0.0f, m3 = 0.0f;
#pragma unroll 2
          for (; a < e_end; a += 16) {
            const uint4 q = lds128u(a);  // one entry: off, w, w*y_x, w*y_y
            const float2 ea = lds64(tab + q.x);
            const float w = __uint_as_float(q.y), wyx = __uint_as_float(q.z), wyy = __uint_as_float(q.w);
            const float X = own.x * ea.x, Y = own.y * ea.y;
            const float Xc = fminf(X, 15.0f), Yc = fminf(Y, 15.0f);
            if (home) {  // warp-uniform
              lp2 = fmaf(wyx, lg2_approx(Xc), lp2);
              lp2 = fmaf(wyy, lg2_approx(Yc), lp2);
              lpw = fmaf(w, Xc + Yc, lpw);
              m1 = fmaxf(m1, Xc);
              m2 = fmaxf(m2, Yc);
              m3 = fmaxf(m3, Xc * Yc);
            }
            gx += X < 15.0f ? fmaf(-w, X, wyx) : 0.0f;
            gy += Y < 15.0f ? fmaf(-w, Y, wyy) : 0.0f;
          }
          if (home) {
            lp_acc += fmaf(lp2, kLn2, -lpw);
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
      }
      ring.release(k);
    }
  }
  const uint32_t b2_0 = __ldg(kp.warp_b2 + vwarp), b2_1 = __ldg(kp.warp_b2 + vwarp + 1);
  ring.begin(kp.stream2 + b2_0, b2_1 - b2_0);  // tau pieces start streaming in during the bounds step

  // ---- bounds (bpl/_util.py:17-31) ------------------------------------------------------------------
  // the maxima and the piece each came from.  One CTA: a 64-bit atomicMax on (value bits | piece offset).  A cluster:
  // 32-bit atomics only (value first, then the owners of the maximum agree on a piece) -- the 64-bit max is a CAS loop
  // for the local CTA and a remote atomic for the others, and the two do not exclude each other (measured lost updates).
  if (S == 1) {
#pragma unroll
    for (int q = 0; q < 3; q++)
      if (best[q] > 0.0f)
        atomicMax(red_best + q * 32 + lane, ((unsigned long long)__float_as_uint(best[q]) << 32) | besth[q]);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const unsigned long long bq = red_best[q * 32 + lane];
      best[q] = __uint_as_float((uint32_t)(bq >> 32));
      besth[q] = (uint32_t)bq;
    }
  } else {
#pragma unroll
    for (int q = 0; q < 3; q++)
      if (best[q] > 0.0f) dsm_atom_max_u32(a_best + (uint32_t)(q * 32 + lane) * 4u, __float_as_uint(best[q]));
    cg::this_cluster().sync();
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const float gq = __uint_as_float(dsm_ld_u32(a_best + (uint32_t)(q * 32 + lane) * 4u));
      if (best[q] == gq && gq > 0.0f) dsm_atom_min_u32(a_best + (uint32_t)((3 + q) * 32 + lane) * 4u, besth[q]);
      best[q] = gq;
    }
    cg::this_cluster().sync();
#pragma unroll
    for (int q = 0; q < 3; q++) besth[q] = dsm_ld_u32(a_best + (uint32_t)((3 + q) * 32 + lane) * 4u);
  }
  const float Lam = fmaxf(best[0], best[1]);
  const int qlam = best[0] >= best[1] ? 0 : 1;
  const float LB = -1.0f / Lam;
  const float UB = fminf(1.0f / best[2], 1.0f);
  const float cc = fmaf(r, UB - LB, LB);

  // ---- arg-max search: virtual warp w looks at entries w, w+VW, ... of each chain's two arg-max pieces -----------
#pragma unroll 1
  for (int which = 0; which < 2; which++) {
    const bool need = (which == 0 || best[2] > 1.0f) && best[which == 0 ? qlam : 2] > 0.0f;  // UB = 1: no dependence on the rates
    const uint32_t hoff = need ? (which == 0 ? (qlam == 0 ? besth[0] : besth[1]) : besth[2]) : b1_0;
    const float target = which == 0 ? Lam : best[2];
    const int q = which == 0 ? qlam : 2;
    const Hdr L = unpack_hdr(__ldg(reinterpret_cast<const uint4*>(kp.stream1 + hoff)));
    const uint32_t n = need ? L.n0 : 0u;
    float2 own = lds64(tab + (need ? L.own_off : 0u));
    if (L.kind >= kH0) { const float s = own.x; own.x = own.y; own.y = s; }
    const unsigned char* ent = kp.stream1 + hoff + 16;
    const uint32_t nmax = __reduce_max_sync(kFull, n);
    uint32_t found = 0xffffffffu, info = 0u;
#pragma unroll 4
    for (uint32_t i = vwarp; i < nmax; i += VW) {
      if (i < n) {
        const uint32_t off = __ldg(reinterpret_cast<const uint32_t*>(ent + (size_t)i * ESZ));
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
          found = (i << 24) | off;
          // own vteam | kind | "X not clipped" | "Y not clipped"
          info = L.vteam | (L.kind << 16) | ((!CLIP || X < 15.0f) ? 1u << 18 : 0u) | ((!CLIP || Y < 15.0f) ? 1u << 19 : 0u);
        }
      }
    }
    // one 64-bit store keeps (entry, piece info) together; several finders only under exact ties inside the piece
    // (then any of them will do: clipped ties carry no gradient)
    if (found != 0xffffffffu)
      dsm_st_u64(a_found + (uint32_t)(which * 32 + lane) * 8u, (unsigned long long)found | ((unsigned long long)info << 32));
  }

  // ---- phase 2: tau terms (bpl/_util.py:54-91) -----------------------------------------------------
  float gc = 0.0f;
  {
    float g[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f}, cacc = 0.0f;
    const uint32_t nst = ring.num_stages();
    for (uint32_t k = 0; k < nst; k++) {
      uint32_t bytes;
      const uint32_t a0 = ring.acquire(k, &bytes);
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
        float lt = 0.0f, uxy = 0.0f, sxy_x = 0.0f, sxy_y = 0.0f;
        {  // tau = 1 - c X Y
          const float Pxy = own.x * own.y;
          const uint32_t e_end = a + L.n0 * (uint32_t)sizeof(Entry);
#pragma unroll 1
          for (; a < e_end; a += 16) {
            const uint4 q = lds128u(a);  // two entries (opponent row offset, w)
#pragma unroll
            for (int j = 0; j < 2; j++) {
              const float2 ea = lds64(tab + (j ? q.z : q.x));
              const float w = __uint_as_float(j ? q.w : q.y);
              if (CLIP) {
                const float Xr = own.x * ea.x, Yr = own.y * ea.y;
                const float t = fminf(Xr, 15.0f) * fminf(Yr, 15.0f);
                const float tau = fmaxf(fmaf(-cc, t, 1.0f), 0.0f);
                const float val = (w * t) * rcp_approx(tau);
                uxy += val;
                sxy_x += Xr < 15.0f ? val : 0.0f;
                sxy_y += Yr < 15.0f ? val : 0.0f;
                if (home) lt = fmaf(w, lg2_approx(tau), lt);
              } else {
                const float t = Pxy * (ea.x * ea.y);
                const float tau = fmaxf(fmaf(-cc, t, 1.0f), 0.0f);
                uxy = fmaf(w * t, rcp_approx(tau), uxy);
                if (home) lt = fmaf(w, lg2_approx(tau), lt);
              }
            }
          }
        }
        float u1[2] = {0.0f, 0.0f}, s1[2] = {0.0f, 0.0f};  // tau = 1 + c X, then tau = 1 + c Y
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const float oc = c == 0 ? own.x : own.y;
          const uint32_t e_end = a + (c == 0 ? L.n1 : L.n2) * (uint32_t)sizeof(Entry);
          float u = 0.0f, sm = 0.0f;
#pragma unroll 1
          for (; a < e_end; a += 16) {
            const uint4 q = lds128u(a);
#pragma unroll
            for (int j = 0; j < 2; j++) {
              const float Rr = oc * lds32(tab + (j ? q.z : q.x));  // `off` already selects .x or .y
              const float w = __uint_as_float(j ? q.w : q.y);
              const float R = CLIP ? fminf(Rr, 15.0f) : Rr;
              const float tau = fmaxf(fmaf(cc, R, 1.0f), 0.0f);
              const float val = (w * R) * rcp_approx(tau);
              u += val;
              if (CLIP) sm += Rr < 15.0f ? val : 0.0f;
              if (home) lt = fmaf(w, lg2_approx(tau), lt);
            }
          }
          u1[c] = u;
          s1[c] = CLIP ? sm : u;
        }
        if (!CLIP) sxy_x = sxy_y = uxy;
        if (home) {
          lp_acc = fmaf(lt, kLn2, lp_acc);
          gc += u1[0] + u1[1] - uxy;
        }
        const float gx = cc * (s1[0] - sxy_x), gy = cc * (s1[1] - sxy_y);
        add_own(g, L.kind, gx, gy);
        if (kp.Cf > 0) {
          cacc += L.kind == kH1 ? gx - gy : gy - gx;
          if (L.flags & kVteamLast) {
            if (ln.active) red_add(ln.sc + (size_t)L.vteam * kp.Cpad, cacc);
            cacc = 0.0f;
          }
        }
        if (L.flags & kTeamLast) put_raw<true>(kp, ln, (int)L.team, g, hacc);
      }
      ring.release(k);
    }
  }
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

  // ---- arg-max fix-up: every warp works out the two matches of its chain and folds
  //      them into the team pass below ---------------------------------------------------------------------
  Fixup fx;
  fx.h1 = 0u;
  fx.confs = 0u;
#pragma unroll
  for (int which = 0; which < 2; which++) {
    const unsigned long long fi = dsm_ld_u64(a_found + (uint32_t)(which * 32 + lane) * 8u);
    const uint32_t packed = (uint32_t)fi;
    fx.teams[which] = 0xffffffffu;
    fx.vts[which] = 0u;
    fx.vx[which] = fx.vy[which] = 0.0f;
    if (packed != 0xffffffffu) {  // else: UB = 1 (or nothing matched: cannot happen, same arithmetic as phase 1)
      const uint32_t info = (uint32_t)(fi >> 32);
      const uint32_t f_off = packed & 0xffffffu;
      const bool h1 = ((info >> 16) & 3u) == kH1;
      const uint32_t own_v = info & 0xffffu;
      const uint32_t opp_v = (f_off - (h1 ? kp.tabQ1 : kp.tabP0)) / kRowBytes;
      const bool xfree = (info >> 18) & 1u, yfree = (info >> 19) & 1u;
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
      fx.teams[which] = (uint32_t)__ldg(kp.v_team + own_v) | ((uint32_t)__ldg(kp.v_team + opp_v) << 16);
      if (kp.Cf > 0)
        fx.confs |= ((uint32_t)__ldg(kp.v_conf + own_v) | ((uint32_t)__ldg(kp.v_conf + opp_v) << 8)) << (16 * which);
    }
  }

  // ---- team pass: raw slots -> parameter gradients, priors, hyper-parameter sums ---------------------------
  const bool has_rho = !dc;
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
#pragma unroll
      for (int which = 0; which < 2; which++) {
        if ((fx.teams[which] & 0xffffu) == (uint32_t)t) fold_fixup(fx, which, 0, ra, rd, rx);
        if ((fx.teams[which] >> 16) == (uint32_t)t) fold_fixup(fx, which, 1, ra, rd, rx);
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
        lp_acc -= 0.5f * (za * za + zd * zd);
        p_za = -za;
        p_zd = -zd;
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
          lp_acc -= 0.5f * dec[i] * dec[i];
          if (ln.active) *ln.g(o.dec[i] + t) = fmaf(hy.sig[i], rx[i], -dec[i]);
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
    float s = __ldg(kp.yconf + k) - cf;
    lp_acc -= 0.5f * cf * cf;
    const int j0 = __ldg(kp.conf_vptr + k), j1 = __ldg(kp.conf_vptr + k + 1);
    for (int j = j0; j < j1; j++) s += ld_cg(ln.sc + (size_t)__ldg(kp.conf_vlist + j) * kp.Cpad);
#pragma unroll
    for (int ws = 0; ws < 4; ws++)
      if (fx.teams[ws >> 1] != 0xffffffffu && ((fx.confs >> (8 * ((ws >> 1) * 2 + (ws & 1)))) & 0xffu) == (uint32_t)k)
        s += fixup_conf(fx, ws >> 1, ws & 1);
    if (ln.active) *ln.g(o.conf + k) = s;
  }

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
      lp += total(0) + 2.0f * (logf(r) + logf(1.0f - r));
      if (ln.active) {
        *ln.g(o.raw) = 2.0f * (1.0f - 2.0f * r) + gc * r * (1.0f - r) * (UB - LB);
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
  red_gc[warp * 32 + lane] = lp;  // every warp read its gc sum before the previous barrier: the rows are free
  __syncthreads();
  if (warp == 0 && ln.active) {
    lp = kp.const_term;
    for (int w = 0; w < W; w++) lp += red_gc[w * 32 + lane];
    kp.lp[chain] = lp;
  }
}

// ---- host launcher --------------------------------------------------------------------------------------
template <bool CLIP>
static int launch_t(const KernelParams& kp, cudaStream_t stream, bool set_attr) {
  auto* fn = &logdensity_kernel<CLIP>;
  if (set_attr) {
    BPLX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return BPLX_OK;
  }
  const int groups = (kp.C + kChains - 1) / kChains;
  if (kp.split > 1) {  // one thread-block cluster per group of chains
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(groups * kp.split));
    cfg.blockDim = dim3((unsigned)(kp.nwarps * 32));
    cfg.dynamicSmemBytes = kp.smem_total;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)kp.split;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    BPLX_CUDA(cudaLaunchKernelEx(&cfg, fn, kp));
  } else {
    fn<<<groups, kp.nwarps * 32, kp.smem_total, stream>>>(kp);
  }
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

int logdensity_set_attributes(const KernelParams& kp) {
  return kp.clip ? launch_t<true>(kp, nullptr, true) : launch_t<false>(kp, nullptr, true);
}
int launch_logdensity(const KernelParams& kp, cudaStream_t stream) {
  return kp.clip ? launch_t<true>(kp, stream, false) : launch_t<false>(kp, stream, false);
}

}  // namespace bplx
