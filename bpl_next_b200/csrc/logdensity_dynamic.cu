// logdensity_dynamic.cu -- K1d: log-density + gradient of the dynamic (random-walk) Dixon-Coles model.
//
// Replaces value_and_grad(potential_fn) of DynamicNeutralDixonColesMatchPredictor._model
// (bpl/dynamic_dixon_coles.py:63-247) with bpl/_util.py:17-93, for a batch of chains; lane = chain.
// Team strengths follow a random walk over gameweeks, attack[j] = attack[j-1] + z[j] * std_attack[j]
// (":192-218"; BPLX_FLAG_DYNAMIC_AS_WRITTEN reproduces the reference literally, where the walk never reaches
// the rates), and a match only pairs teams of its own gameweek.  So:
//   prefix pass   (team-owned)      the walk's prefix sums of every (gameweek, team) -> workspace
//   phase 1 / 2   (gameweek-owned)  a warp owns whole gameweeks; at a gameweek's marker piece it rebuilds the
//                                   tables of that gameweek in its private slice of shared memory, then walks the
//                                   gameweek's list pieces exactly like K1 (logdensity.cu).  The arg-max search
//                                   runs in phase 2 while the arg-max piece's gameweek is resident.
//   suffix pass   (team-owned)      d/d attack[j] summed over the later gameweeks (the walk's transpose)
//   gameweek pass (gameweek-owned)  priors, chain rule and the ten per-gameweek hyper-parameter gradients,
//                                   which need no cross-warp reduction because one warp sees the whole gameweek.
#include <math.h>

#include "k1_common.cuh"
#include "problem.h"

namespace bplx {

__global__ void __launch_bounds__(kMaxWarpsDyn * 32, 1) logdensity_dynamic_kernel(const __grid_constant__ KernelParams kp) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = kp.nwarps;
  const int T = kp.T, G = kp.G;
  const ThetaOffsets& o = kp.off;
  const int chain_raw = blockIdx.x * kChains + lane;
  const int chain = min(chain_raw, kp.C - 1);
  Lane ln;
  ln.th = kp.theta + (size_t)chain * (size_t)kp.sc;
  ln.gr = kp.grad + (size_t)chain * (size_t)kp.sc;
  ln.sc = kp.scratch + chain;  // prefix sums: [(j*T + t)*2 + {att, def}][Cpad]
  ln.sd = kp.sd;
  ln.active = chain_raw < kp.C;
  const uint32_t tab = smem_u32(smem) + warp * kp.tab_bytes + lane * 8;  // this warp's private tables
  unsigned long long* red_best = reinterpret_cast<unsigned long long*>(smem + kp.smem_red);  // [3][32]
  uint32_t* red_found = reinterpret_cast<uint32_t*>(smem + kp.smem_red + 768);               // [2][32]
  uint32_t* red_info = reinterpret_cast<uint32_t*>(smem + kp.smem_red + 1024);               // [2][32]
  float* red_gc = reinterpret_cast<float*>(smem + kp.smem_red + 1280 + 12 * 128);             // [W][32]
  constexpr uint32_t ESZ = (uint32_t)sizeof(Entry);
  Ring ring;
  ring.init(smem_u32(smem) + kp.smem_ring + warp * (kStages * kp.stage_bytes),
            smem_u32(smem) + kp.smem_bar + warp * (kStages * 8), kp.stage_bytes, lane);
  const uint32_t b1_0 = __ldg(kp.warp_b1 + warp), b1_1 = __ldg(kp.warp_b1 + warp + 1);
  ring.begin(kp.stream1 + b1_0, b1_1 - b1_0);
  {  // pull this CTA's slice of theta into L2
    const int nthr = W * 32;
    if (kp.sd == 1) {
      const int per = (kp.D + 31) / 32;
      for (int i = threadIdx.x; i < 32 * per; i += nthr) {
        const int c = min(blockIdx.x * kChains + i / per, kp.C - 1);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(kp.theta + (size_t)c * (size_t)kp.sc + (size_t)(i % per) * 32));
      }
    } else {
      for (int d = threadIdx.x; d < kp.D; d += nthr)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(kp.theta + (size_t)d * (size_t)kp.sd + (size_t)blockIdx.x * kChains));
    }
  }
  float lp_acc = 0.0f;
  const float r = sigmoid_clipped(ln.ld(o.raw));
  const float mu_d = ln.ld(o.mean_defence);

  // ---- prefix pass: attack / defence of every (gameweek, team) ----------------------------------------
  {
    const uint32_t zr = (uint32_t)T * kRowBytes;  // zero rows of the private tables
    if (kp.has1) {
      sts64(tab + kp.tabP1 + zr, 0.0f, 0.0f);
      sts64(tab + kp.tabQ1 + zr, 0.0f, 0.0f);
    }
    if (kp.has0) sts64(tab + kp.tabP0 + zr, 0.0f, 0.0f);
    if (warp == 0) {
      red_best[lane] = red_best[32 + lane] = red_best[64 + lane] = 0ull;
      red_found[lane] = red_found[32 + lane] = 0xffffffffu;
    }
    for (int t = warp; t < T; t += W) {
      float att = 0.0f, def = mu_d;
      for (int k = 0; k < kp.K; k++) {
        const float x = __ldg(kp.Xs + (size_t)t * kp.K + k);
        att = fmaf(x, ln.ld(o.beta_a + k), att);
        def = fmaf(x, ln.ld(o.beta_d + k), def);
      }
      for (int j = 0; j < G; j++) {
        const int jt = j * T + t;
        if (kp.as_written) {
          att = def = 0.0f;  // dynamic_dixon_coles.py:192-218: the .at[].set results are discarded
        } else {
          att = fmaf(ln.ld(o.za + jt), expf(ln.ld(o.log_std_attack + j)), att);
          def = fmaf(ln.ld(o.zd + jt), expf(ln.ld(o.log_std_defence + j)), def);
        }
        if (ln.active) {
          ln.sc[(size_t)(2 * jt) * kp.Cpad] = att;
          ln.sc[(size_t)(2 * jt + 1) * kp.Cpad] = def;
        }
      }
    }
  }
  __syncthreads();

  // tables of gameweek j in this warp's slice; with_lp: add the static sum of w * y * log(lambda)
  auto build_tables = [&](int j, bool with_lp) {
    float mu[4], sig[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      mu[i] = ln.ld(o.mean[i] + j);
      sig[i] = expf(ln.ld(o.log_std[i] + j));
    }
    const int i0 = __ldg(kp.gw_tptr + j), i1 = __ldg(kp.gw_tptr + j + 1);
    constexpr int B = 4;  // teams per batch: all loads of a batch are in flight together
    for (int ib = i0; ib < i1; ib += B) {
      int tt[B];
      float att[B], def[B], dec[B][4];
#pragma unroll
      for (int b = 0; b < B; b++) {
        tt[b] = __ldg(kp.gw_tlist + min(ib + b, i1 - 1));
        const int jt = j * T + tt[b];
        att[b] = ld_cg(ln.sc + (size_t)(2 * jt) * kp.Cpad);
        def[b] = ld_cg(ln.sc + (size_t)(2 * jt + 1) * kp.Cpad);
#pragma unroll
        for (int i = 0; i < 4; i++) dec[b][i] = ln.ld(o.dec[i] + jt);
      }
#pragma unroll
      for (int b = 0; b < B; b++) {
        if (ib + b >= i1) break;
        const int t = tt[b], jt = j * T + t;
        float x[4];
#pragma unroll
        for (int i = 0; i < 4; i++) x[i] = fmaf(sig[i], dec[b][i], mu[i]);
        float ex[6];
        ex[eAh1] = att[b] + x[0];
        ex[eBh1] = -def[b] - x[2];
        ex[eBa1] = -def[b] - x[3];
        ex[eAa1] = att[b] + x[1];
        ex[eA0] = att[b];
        ex[eB0] = -def[b];
        const uint32_t row = (uint32_t)t * kRowBytes;
        if (kp.has1) {
          sts64(tab + kp.tabP1 + row, expf(ex[eAh1]), expf(ex[eBh1]));
          sts64(tab + kp.tabQ1 + row, expf(ex[eBa1]), expf(ex[eAa1]));
        }
        if (kp.has0) sts64(tab + kp.tabP0 + row, expf(ex[eA0]), expf(ex[eB0]));
        if (with_lp) {
#pragma unroll
          for (int e = 0; e < 6; e++) lp_acc = fmaf(__ldg(kp.yexp + (size_t)jt * 6 + e), ex[e], lp_acc);
        }
      }
    }
  };
  auto slot = [&](int base, int jt) { return ln.g(base + jt); };

  // ---- phase 1 ----------------------------------------------------------------------------------
  float best[3] = {0.0f, 0.0f, 0.0f};
  uint32_t besth[3] = {0u, 0u, 0u};
  {
    float g[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    const uint32_t nst = ring.num_stages();
    for (uint32_t k = 0; k < nst; k++) {
      uint32_t bytes;
      const uint32_t a0 = ring.acquire(k, &bytes);
      uint32_t a = a0;
      const uint32_t aend = a0 + bytes;
      while (a + 16 <= aend) {
        const uint32_t hoff = b1_0 + k * ring.S + (a - a0);
        const Hdr L = unpack_hdr(lds128u(a));
        a += 16;
        if (L.flags & kGwFirst) build_tables((int)L.vteam, true);
        if (L.team == 0xffffu) continue;  // marker or filler
        const uint32_t e_end = a + L.n0 * ESZ;
        if (L.flags & kTeamFirst) {
#pragma unroll
          for (int e = 0; e < 6; e++) g[e] = 0.0f;
        }
        float2 own = lds64(tab + L.own_off);
        if (L.kind >= kH0) { const float s = own.x; own.x = own.y; own.y = s; }
        const bool home = (L.kind & 1) == 0;
        float ax0 = 0.0f, ay0 = 0.0f, ax1 = 0.0f, ay1 = 0.0f, m1 = 0.0f, m2 = 0.0f, m3 = 0.0f;
#pragma unroll 4
        for (; a < e_end; a += 16) {
          const uint4 q = lds128u(a);  // two entries
          const float2 ea = lds64(tab + q.x), eb = lds64(tab + q.z);
          const float wa = __uint_as_float(q.y), wb = __uint_as_float(q.w);
          ax0 = fmaf(wa, ea.x, ax0); ay0 = fmaf(wa, ea.y, ay0);
          ax1 = fmaf(wb, eb.x, ax1); ay1 = fmaf(wb, eb.y, ay1);
          if (home) {  // warp-uniform
            m1 = fmaxf(m1, fmaxf(ea.x, eb.x));
            m2 = fmaxf(m2, fmaxf(ea.y, eb.y));
            m3 = fmaxf(m3, fmaxf(ea.x * ea.y, eb.x * eb.y));
          }
        }
        if (home) {
          const float v0 = own.x * m1, v1 = own.y * m2, v2 = (own.x * own.y) * m3;
          if (v0 > best[0]) { best[0] = v0; besth[0] = hoff; }
          if (v1 > best[1]) { best[1] = v1; besth[1] = hoff; }
          if (v2 > best[2]) { best[2] = v2; besth[2] = hoff; }
        }
        const float SX = own.x * (ax0 + ax1), SY = own.y * (ay0 + ay1);
        lp_acc -= 0.5f * (SX + SY);  // every match is in two lists
        add_own(g, L.kind, -SX, -SY);
        if ((L.flags & kTeamLast) && ln.active) {
          const int jt = (int)L.vteam * T + (int)L.team;
          *slot(o.za, jt) = g[eAh1] + g[eAa1] + g[eA0];
          *slot(o.zd, jt) = -(g[eBh1] + g[eBa1] + g[eB0]);
          *slot(o.dec[0], jt) = g[eAh1];
          *slot(o.dec[1], jt) = g[eAa1];
          *slot(o.dec[2], jt) = -g[eBh1];
          *slot(o.dec[3], jt) = -g[eBa1];
        }
      }
      ring.release(k);
    }
  }
  const uint32_t b2_0 = __ldg(kp.warp_b2 + warp), b2_1 = __ldg(kp.warp_b2 + warp + 1);
  ring.begin(kp.stream2 + b2_0, b2_1 - b2_0);

  // ---- bounds (bpl/_util.py:17-31) ------------------------------------------------------------------
#pragma unroll
  for (int q = 0; q < 3; q++)
    if (best[q] > 0.0f)
      atomicMax(red_best + q * 32 + lane, ((unsigned long long)__float_as_uint(best[q]) << 32) | besth[q]);
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 3; q++) {
    const unsigned long long b = red_best[q * 32 + lane];
    best[q] = __uint_as_float((uint32_t)(b >> 32));
    besth[q] = (uint32_t)b;
  }
  const float Lam = fmaxf(best[0], best[1]);
  const int qlam = best[0] >= best[1] ? 0 : 1;
  const float LB = -1.0f / Lam;
  const float UB = fminf(1.0f / best[2], 1.0f);
  const float cc = fmaf(r, UB - LB, LB);
  const uint32_t hoff0 = qlam == 0 ? besth[0] : besth[1], hoff1 = besth[2];
  // gameweeks of the two arg-max pieces (0xffff = no search needed)
  uint32_t gwsel = unpack_hdr(__ldg(reinterpret_cast<const uint4*>(kp.stream1 + hoff0))).vteam;
  gwsel |= (best[2] > 1.0f ? unpack_hdr(__ldg(reinterpret_cast<const uint4*>(kp.stream1 + hoff1))).vteam : 0xffffu) << 16;

  // ---- phase 2: tau terms (bpl/_util.py:54-91) + arg-max search in the resident gameweek --------------
  float gc = 0.0f;
  {
    float g[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    const uint32_t nst = ring.num_stages();
    for (uint32_t k = 0; k < nst; k++) {
      uint32_t bytes;
      const uint32_t a0 = ring.acquire(k, &bytes);
      uint32_t a = a0;
      const uint32_t aend = a0 + bytes;
      while (a + 16 <= aend) {
        const Hdr L = unpack_hdr(lds128u(a));
        a += 16;
        if (L.flags & kGwFirst) {
          const uint32_t j = L.vteam;
          build_tables((int)j, false);
#pragma unroll 1
          for (int which = 0; which < 2; which++) {
            const bool need = ((gwsel >> (16 * which)) & 0xffffu) == j;
            if (!__any_sync(kFull, need)) continue;
            const uint32_t hoff = which == 0 ? hoff0 : hoff1;
            const float target = which == 0 ? Lam : best[2];
            const int q = which == 0 ? qlam : 2;
            const Hdr P = unpack_hdr(__ldg(reinterpret_cast<const uint4*>(kp.stream1 + hoff)));
            const uint32_t n = need ? P.n0 : 0u;
            float2 own = lds64(tab + (need ? P.own_off : 0u));
            if (P.kind >= kH0) { const float s = own.x; own.x = own.y; own.y = s; }
            const unsigned char* ent = kp.stream1 + hoff + 16;
            const uint32_t nmax = __reduce_max_sync(kFull, n);
            uint32_t found = 0xffffffffu;
            for (uint32_t i = 0; i < nmax; i++) {
              if (i < n && found == 0xffffffffu) {
                const uint32_t off = __ldg(reinterpret_cast<const uint32_t*>(ent + (size_t)i * ESZ));
                const float2 ea = lds64(tab + off);
                const float val = q == 0 ? own.x * ea.x : (q == 1 ? own.y * ea.y : (own.x * own.y) * (ea.x * ea.y));
                if (val == target) found = (i << 24) | off;
              }
            }
            if (found != 0xffffffffu) {
              red_found[which * 32 + lane] = found;
              red_info[which * 32 + lane] = P.team | (P.kind << 16) | (3u << 18);  // own team | kind | nothing clipped
            }
          }
        }
        if (L.team == 0xffffu) continue;
        if (L.flags & kTeamFirst) {
#pragma unroll
          for (int e = 0; e < 6; e++) g[e] = 0.0f;
        }
        float2 own = lds64(tab + L.own_off);
        if (L.kind >= kH0) { const float s = own.x; own.x = own.y; own.y = s; }
        const bool home = (L.kind & 1) == 0;
        float lt = 0.0f, uxy = 0.0f;
        {  // tau = 1 - c X Y
          const float Pxy = own.x * own.y;
          const uint32_t e_end = a + L.n0 * ESZ;
#pragma unroll 1
          for (; a < e_end; a += 16) {
            const uint4 q = lds128u(a);
#pragma unroll
            for (int jj = 0; jj < 2; jj++) {
              const float2 ea = lds64(tab + (jj ? q.z : q.x));
              const float w = __uint_as_float(jj ? q.w : q.y);
              const float t = Pxy * (ea.x * ea.y);
              const float tau = fmaxf(fmaf(-cc, t, 1.0f), 0.0f);
              uxy = fmaf(w * t, rcp_approx(tau), uxy);
              if (home) lt = fmaf(w, lg2_approx(tau), lt);
            }
          }
        }
        float u1[2] = {0.0f, 0.0f};  // tau = 1 + c X, then tau = 1 + c Y
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const float oc = c == 0 ? own.x : own.y;
          const uint32_t e_end = a + (c == 0 ? L.n1 : L.n2) * ESZ;
          float u = 0.0f;
#pragma unroll 1
          for (; a < e_end; a += 16) {
            const uint4 q = lds128u(a);
#pragma unroll
            for (int jj = 0; jj < 2; jj++) {
              const float R = oc * lds32(tab + (jj ? q.z : q.x));
              const float w = __uint_as_float(jj ? q.w : q.y);
              const float tau = fmaxf(fmaf(cc, R, 1.0f), 0.0f);
              u = fmaf(w * R, rcp_approx(tau), u);
              if (home) lt = fmaf(w, lg2_approx(tau), lt);
            }
          }
          u1[c] = u;
        }
        if (home) {
          lp_acc = fmaf(lt, kLn2, lp_acc);
          gc += u1[0] + u1[1] - uxy;
        }
        add_own(g, L.kind, cc * (u1[0] - uxy), cc * (u1[1] - uxy));
        if ((L.flags & kTeamLast) && ln.active) {
          const int jt = (int)L.vteam * T + (int)L.team;
          red_add(slot(o.za, jt), g[eAh1] + g[eAa1] + g[eA0]);
          red_add(slot(o.zd, jt), -(g[eBh1] + g[eBa1] + g[eB0]));
          red_add(slot(o.dec[0], jt), g[eAh1]);
          red_add(slot(o.dec[1], jt), g[eAa1]);
          red_add(slot(o.dec[2], jt), -g[eBh1]);
          red_add(slot(o.dec[3], jt), -g[eBa1]);
        }
      }
      ring.release(k);
    }
  }
  red_gc[warp * 32 + lane] = gc;
  __syncthreads();  // tables are dead from here on; raw slots hold both phases
  gc = 0.0f;
  for (int w = 0; w < W; w++) gc += red_gc[w * 32 + lane];
  {  // the 1-1 matches: tau = 1 - c for all of them
    const float t11 = fmaxf(1.0f - cc, 0.0f);
    gc -= kp.w11 / t11;
    if (warp == 0 && kp.w11 != 0.0f) lp_acc = fmaf(kp.w11, logf(t11), lp_acc);
  }

  // ---- arg-max fix-up descriptors ---------------------------------------------------
  Fixup fx;
  fx.h1 = 0u;
  fx.confs = 0u;
#pragma unroll
  for (int which = 0; which < 2; which++) {
    const uint32_t packed = red_found[which * 32 + lane];
    fx.teams[which] = 0xffffffffu;
    fx.vts[which] = 0xffffu;  // the gameweek
    fx.vx[which] = fx.vy[which] = 0.0f;
    if (packed != 0xffffffffu) {
      const uint32_t info = red_info[which * 32 + lane];
      const uint32_t f_off = packed & 0xffffffu;
      const bool h1 = ((info >> 16) & 3u) == kH1;
      const uint32_t opp_t = ((f_off & ~7u) - (h1 ? kp.tabQ1 : kp.tabP0)) / kRowBytes;
      if (which == 0) {
        const float wgt = gc * (1.0f - r) / Lam;  // dc/dLB * dLB/d eta
        fx.vx[0] = qlam == 0 ? wgt : 0.0f;
        fx.vy[0] = qlam == 1 ? wgt : 0.0f;
      } else {
        const float wgt = -gc * r / best[2];  // dc/dUB * dUB/d eta
        fx.vx[1] = fx.vy[1] = wgt;
      }
      fx.h1 |= (h1 ? 1u : 0u) << which;
      fx.teams[which] = (info & 0xffffu) | (opp_t << 16);
      fx.vts[which] = (gwsel >> (16 * which)) & 0xffffu;
    }
  }

  // ---- suffix pass (team-owned): complete the raw slots; attack / defence slots <- sums over later gameweeks --
  for (int t = warp; t < T; t += W) {
    float s_att = 0.0f, s_def = 0.0f;
    constexpr int B = 2;  // gameweeks per batch: all loads of a batch are in flight together
    for (int jb = G - 1; jb >= 0; jb -= B) {
      float ra[B], rd[B], rx[B][4];
#pragma unroll
      for (int b = 0; b < B; b++) {
        const int j = max(jb - b, 0), jt = j * T + t;
        const float4 ys = __ldg(reinterpret_cast<const float4*>(kp.yteam + (size_t)jt * 8));
        const float2 ys2 = __ldg(reinterpret_cast<const float2*>(kp.yteam + (size_t)jt * 8 + 4));
        ra[b] = ys.x; rd[b] = ys.y; rx[b][0] = ys.z; rx[b][1] = ys.w; rx[b][2] = ys2.x; rx[b][3] = ys2.y;
        if (__ldg(kp.team_flags + jt) & 1) {  // phase 1 wrote the slots of this (gameweek, team)
          ra[b] += ld_cg(slot(o.za, jt));
          rd[b] += ld_cg(slot(o.zd, jt));
#pragma unroll
          for (int i = 0; i < 4; i++) rx[b][i] += ld_cg(slot(o.dec[i], jt));
        }
      }
#pragma unroll
      for (int b = 0; b < B; b++) {
        const int j = jb - b;
        if (j < 0) break;
        const int jt = j * T + t;
#pragma unroll
        for (int which = 0; which < 2; which++) {
          if (fx.vts[which] == (uint32_t)j) {
            if ((fx.teams[which] & 0xffffu) == (uint32_t)t) fold_fixup(fx, which, 0, ra[b], rd[b], rx[b]);
            if ((fx.teams[which] >> 16) == (uint32_t)t) fold_fixup(fx, which, 1, ra[b], rd[b], rx[b]);
          }
        }
        s_att += ra[b];
        s_def += rd[b];
        if (ln.active) {
          *slot(o.za, jt) = kp.as_written ? 0.0f : s_att;
          *slot(o.zd, jt) = kp.as_written ? 0.0f : s_def;
#pragma unroll
          for (int i = 0; i < 4; i++) *slot(o.dec[i], jt) = rx[b][i];
        }
      }
    }
  }
  __syncthreads();

  // ---- gameweek pass (gameweek-owned): priors, chain rule, per-gameweek hyper-parameter gradients ------------
  float lp = 0.0f;  // this warp's part of the prior terms
  for (int j = warp; j < G; j += W) {
    if (j == 0) {  // mean_defence and the covariate coefficients see the whole walk: sum_t d/d attack[0], defence[0]
      float s = 0.0f;
      for (int t = 0; t < T; t++) s += ld_cg(slot(o.zd, t));
      lp -= 0.5f * mu_d * mu_d;
      if (ln.active) *ln.g(o.mean_defence) = s - mu_d;
      for (int k = 0; k < kp.K; k++) {
        float sa = 0.0f, sd = 0.0f;
        for (int t = 0; t < T; t++) {
          const float x = __ldg(kp.Xs + (size_t)t * kp.K + k);
          sa = fmaf(x, ld_cg(slot(o.za, t)), sa);
          sd = fmaf(x, ld_cg(slot(o.zd, t)), sd);
        }
        const float ba = ln.ld(o.beta_a + k), bd = ln.ld(o.beta_d + k);
        lp -= 0.5f * (ba * ba + bd * bd);
        if (ln.active) {
          *ln.g(o.beta_a + k) = sa - ba;
          *ln.g(o.beta_d + k) = sd - bd;
        }
      }
    }
    const float lsa = ln.ld(o.log_std_attack + j), lsd = ln.ld(o.log_std_defence + j);
    const float sig_a = expf(lsa), sig_d = expf(lsd);
    float mu[4], lsig[4], sig[4], a_mu[4], a_ls[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      mu[i] = ln.ld(o.mean[i] + j);
      lsig[i] = ln.ld(o.log_std[i] + j);
      sig[i] = expf(lsig[i]);
      a_mu[i] = a_ls[i] = 0.0f;
    }
    float a_ls_a = 0.0f, a_ls_d = 0.0f;
    constexpr int B = 2;  // teams per batch: all loads of a batch are in flight together
    for (int tb = 0; tb < T; tb += B) {
      float za[B], zd[B], ul[B], s_att[B], s_def[B], dec[B][4], rx[B][4];
#pragma unroll
      for (int b = 0; b < B; b++) {
        const int jt = j * T + min(tb + b, T - 1);
        za[b] = ln.ld(o.za + jt); zd[b] = ln.ld(o.zd + jt); ul[b] = ln.ld(o.u + jt);
        s_att[b] = ld_cg(slot(o.za, jt)); s_def[b] = ld_cg(slot(o.zd, jt));
#pragma unroll
        for (int i = 0; i < 4; i++) {
          dec[b][i] = ln.ld(o.dec[i] + jt);
          rx[b][i] = ld_cg(slot(o.dec[i], jt));
        }
      }
#pragma unroll
      for (int b = 0; b < B; b++) {
        if (tb + b >= T) break;
        const int jt = j * T + tb + b;
        const float u = sigmoid_clipped(ul[b]);
        const float rho = 2.0f * u - 1.0f, inv_s2 = 1.0f / (1.0f - rho * rho);
        const float e = zd[b] - rho * za[b], es = e * inv_s2;
        // u ~ Beta(2,4) + Jacobian; za ~ N(0,1); zd ~ N(rho za, sqrt(1 - rho^2))  (dynamic_dixon_coles.py:128-143)
        lp += -0.5f * (za[b] * za[b] + e * es) + 0.5f * logf(inv_s2) + 2.0f * logf(u) + 4.0f * logf(1.0f - u);
        const float a_rho = es * za[b] - rho * es * es + rho * inv_s2;
        if (ln.active) {
          *slot(o.u, jt) = 2.0f - 6.0f * u + a_rho * 2.0f * u * (1.0f - u);
          *slot(o.za, jt) = fmaf(sig_a, s_att[b], -za[b] + rho * es);
          *slot(o.zd, jt) = fmaf(sig_d, s_def[b], -es);
        }
        a_ls_a = fmaf(sig_a * za[b], s_att[b], a_ls_a);
        a_ls_d = fmaf(sig_d * zd[b], s_def[b], a_ls_d);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          lp -= 0.5f * dec[b][i] * dec[b][i];
          if (ln.active) *slot(o.dec[i], jt) = fmaf(sig[i], rx[b][i], -dec[b][i]);
          a_mu[i] += rx[b][i];
          a_ls[i] = fmaf(sig[i] * dec[b][i], rx[b][i], a_ls[i]);
        }
      }
    }
    // the ten hyper-parameters of gameweek j (dynamic_dixon_coles.py:74-98)
    lp += fmaf(-0.5f * sig_a, sig_a, lsa) + fmaf(-0.5f * sig_d, sig_d, lsd);
    if (ln.active) {
      *ln.g(o.log_std_attack + j) = fmaf(-sig_a, sig_a, 1.0f) + a_ls_a;
      *ln.g(o.log_std_defence + j) = fmaf(-sig_d, sig_d, 1.0f) + a_ls_d;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float z = (mu[i] - ((i & 1) ? -0.1f : 0.1f)) * 5.0f;  // N(+-0.1, 0.2)
      lp += -0.5f * z * z + fmaf(-0.5f * sig[i], sig[i], lsig[i]);
      if (ln.active) {
        *ln.g(o.mean[i] + j) = fmaf(-z, 5.0f, a_mu[i]);
        *ln.g(o.log_std[i] + j) = fmaf(-sig[i], sig[i], 1.0f) + a_ls[i];
      }
    }
  }
  if (warp == (W > 1 ? 1 : 0)) {  // corr_coef_raw ~ Uniform(0,1): Jacobian only; corr_coef = LB + r (UB - LB)
    lp += logf(r) + logf(1.0f - r);
    if (ln.active) {
      *ln.g(o.raw) = (1.0f - 2.0f * r) + gc * r * (1.0f - r) * (UB - LB);
      if (kp.corr_coef) kp.corr_coef[chain] = cc;
    }
  }
  red_gc[warp * 32 + lane] = lp + lp_acc;  // every warp read its gc sum two barriers ago: the rows are free
  __syncthreads();
  if (warp == 0 && ln.active) {
    lp = kp.const_term;
    for (int w = 0; w < W; w++) lp += red_gc[w * 32 + lane];
    // a log-density is never +inf: that is an intermediate that overflowed float32 far from the typical set (a sampler
    // would accept such a point as the best ever seen); NaN is what the callers reject
    kp.lp[chain] = lp == INFINITY ? NAN : lp;
  }
}

int logdensity_dynamic_set_attributes() {
  BPLX_CUDA(cudaFuncSetAttribute(&logdensity_dynamic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return BPLX_OK;
}
int launch_logdensity_dynamic(const KernelParams& kp, cudaStream_t stream) {
  const int grid = (kp.C + kChains - 1) / kChains;
  logdensity_dynamic_kernel<<<grid, kp.nwarps * 32, kp.smem_total, stream>>>(kp);
  BPLX_CUDA(cudaGetLastError());
  note_launch(1);
  return BPLX_OK;
}

}  // namespace bplx
