// logdensity_few.cu -- K1s: the few-chain ("streaming") regime of the log-density + gradient, one CTA per CHAIN.
//
// Same model arithmetic as K1 (logdensity.cu; bpl/dixon_coles.py:39-84, bpl/extended_dixon_coles.py:78-248,
// bpl/neutral_dixon_coles.py:102-283, bpl/neutral_dixon_coles_WC.py:83-232, bpl/_util.py:17-93), same static plan
// (plan.h: the split-1 streams of list pieces), but the parallel axis is the MATCHES of one chain instead of the chains:
// K1 gives every chain a lane and needs 32 chains per SM before a lane is busy; with a handful of chains (the reference's
// own default is ONE chain, bpl/dixon_coles.py:101-106) it leaves the GPU idle and a call costs the whole plan walked by
// one group of lanes (configs[2] data, 1-32 chains: 65 us even with an 8-CTA cluster per group).  Here
//
//   tables     one float2 per (table, virtual team) in shared memory (K1's table area / 32)
//   phase 1    a warp walks the same per-warp stream of pieces as in K1; the LANES take the entries of a piece (one
//              coalesced 256-byte load per 32 entries), sums and maxima meet in xor-shuffle trees (fixed order: deterministic);
//              the arg-max ENTRY travels with the maximum, so no search pass is needed
//   bounds     maxima over warps -> LB, UB, corr_coef
//   phase 2    tau pieces the same way
//   team pass  one THREAD per team: raw slots (shared memory) -> parameter gradients, priors, hyper sums (block reduction)
//   epilogue   scalar sites by single threads
// Entered through the same bplx_logdensity_fwdbwd / bplx_loglik_fwdbwd calls: api.cu picks it for small chain counts.
#include <math.h>

#include "k1_common.cuh"
#include "problem.h"

namespace bplx {

namespace {

__device__ __forceinline__ uint2 lds64u(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
// maximum with the lowest index among equals (the same in every lane afterwards)
__device__ __forceinline__ void wargmax(float& v, uint32_t& idx) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ov = __shfl_xor_sync(kFull, v, o);
    const uint32_t oi = __shfl_xor_sync(kFull, idx, o);
    if (ov > v || (ov == v && oi < idx)) {
      v = ov;
      idx = oi;
    }
  }
}

struct FewSmem {
  float2* tabs;   // [rows] exponentials, row = K1 byte offset >> 8
  float* raw1;    // [T][8]   phase-1 slots: d/d att, d/d def, d/d venue effect[4]
  float* raw2;    // [2][T][8] phase-2 slots by side of the run's last piece
  float* conf1;   // [V]      phase-1 d/d (A - B) per virtual team
  float* conf2;   // [2][V]
  float* rows;    // [T][2]   team rows for the covariate coefficients
  float* red;     // [W][16]  block reductions
  float* tot;     // [16]     their totals
  float* bestv;   // [W][3]
  uint32_t* besti;  // [W][3]  piece offset << 8 | entry
  float* misc;    // [8]
};

}  // namespace

template <bool CLIP>
__global__ void __launch_bounds__(kMaxWarps * 32) logdensity_few_kernel(const __grid_constant__ KernelParams kp,
                                                                        const __grid_constant__ WarpBounds wb) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = kp.nwarps, NT = W * 32;
  const int T = kp.T, V = kp.V, ndec = kp.ndec;
  const ThetaOffsets& o = kp.off;
  const int chain = blockIdx.x;
  Lane ln;
  ln.th = kp.theta + (size_t)chain * (size_t)kp.sc;
  ln.gr = kp.grad + (size_t)chain * (size_t)kp.sc;
  ln.sc = nullptr;
  ln.sd = kp.sd;
  ln.active = true;
  const int nrows = (int)(kp.tab_bytes >> 8) + 1;
  FewSmem S;
  const uint32_t ring_bytes = (uint32_t)W * kStages * kp.stage_bytes;
  Ring ring;  // the warp's TMA ring over its stream of pieces, as in K1
  ring.init(smem_u32(smem) + warp * (kStages * kp.stage_bytes), smem_u32(smem) + ring_bytes + warp * (kStages * 8),
            kp.stage_bytes, lane);
  const uint32_t b1_0 = wb.b1[warp], b1_1 = wb.b1[warp + 1];
  ring.begin(kp.stream1 + b1_0, b1_1 - b1_0);  // phase-1 pieces stream in while the prologue runs
  {
    unsigned char* p = smem + ring_bytes + (size_t)W * kStages * 8;
    S.tabs = reinterpret_cast<float2*>(p); p += (size_t)nrows * 8;
    S.raw1 = reinterpret_cast<float*>(p); p += (size_t)T * 32;
    S.raw2 = reinterpret_cast<float*>(p); p += (size_t)T * 64;
    S.conf1 = reinterpret_cast<float*>(p); p += (size_t)(V + 1) * 4;
    S.conf2 = reinterpret_cast<float*>(p); p += (size_t)(V + 1) * 8;
    S.rows = reinterpret_cast<float*>(p); p += (size_t)T * 8;
    S.red = reinterpret_cast<float*>(p); p += (size_t)W * 64;
    S.tot = reinterpret_cast<float*>(p); p += 64;
    S.bestv = reinterpret_cast<float*>(p); p += (size_t)W * 12;
    S.besti = reinterpret_cast<uint32_t*>(p); p += (size_t)W * 12;
    S.misc = reinterpret_cast<float*>(p);
  }
  // deterministic block sums of up to 16 values per thread: warp trees, then the warps in order
  auto block_sum = [&](const float* v, int n) {
    __syncthreads();  // the previous use of red / tot is over
    for (int i = 0; i < n; i++) {
      const float s = wsum(v[i]);
      if (lane == 0) S.red[warp * 16 + i] = s;
    }
    __syncthreads();
    if (tid < n) {
      float s = 0.0f;
      for (int w = 0; w < W; w++) s += S.red[w * 16 + tid];
      S.tot[tid] = s;
    }
    __syncthreads();
  };

  const bool dc = ndec == 0, lik = kp.lik_only != 0;
  const bool has_rho = !dc && !lik;
  const float pri = lik ? 0.0f : 1.0f;
  const Hyp hy = load_hyp(kp, ln);
  const float raw_in = ln.ld(o.raw);
  const float r = lik ? raw_in : sigmoid_clipped(raw_in);
  const float u = has_rho ? sigmoid_clipped(ln.ld(o.u)) : 0.5f;
  for (int i = tid; i < T * 8; i += NT) S.raw1[i] = 0.0f;
  for (int i = tid; i < T * 16; i += NT) S.raw2[i] = 0.0f;
  for (int i = tid; i < 3 * (V + 1); i += NT) S.conf1[i] = 0.0f;  // (conf1 and conf2 are contiguous)
  for (int i = tid; i < nrows; i += NT) S.tabs[i] = make_float2(0.0f, 0.0f);  // zero rows included
  __syncthreads();

  float lp_acc = 0.0f;  // this thread's share of the log-density
  float hacc = 0.0f;    // DIXON_COLES: d/d home_advantage (warp-uniform, counted by lane 0)
  // ---- prologue: one thread per team ---------------------------------------------------------------------
  for (int t = tid; t < T; t += NT) {
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
      if (kp.has1) {
        S.tabs[(kp.tabP1 >> 8) + v] = make_float2(expf(ex[eAh1]), expf(ex[eBh1]));
        S.tabs[(kp.tabQ1 >> 8) + v] = make_float2(expf(ex[eBa1]), expf(ex[eAa1]));
      }
      if (kp.has0) S.tabs[(kp.tabP0 >> 8) + v] = make_float2(expf(ex[eA0]), expf(ex[eB0]));
      if (!CLIP) {  // static sum of w * y * log(lambda): linear in the exponents
#pragma unroll
        for (int e = 0; e < 6; e++) lp_acc = fmaf(__ldg(kp.yexp + (size_t)v * 6 + e), ex[e], lp_acc);
      }
    }
  }
  __syncthreads();

  auto row = [&](uint32_t off) { return S.tabs[off >> 8]; };
  constexpr uint32_t ESZ = CLIP ? (uint32_t)sizeof(EntryClip) : (uint32_t)sizeof(Entry);
  // ---- phase 1: the warp's stream of rate pieces, lanes over the entries ------------------------------------
  // every lane keeps the largest rate / product it has seen with the piece and entry it came from; they meet once, at
  // the end of the stream (a rate is own factor x opponent factor: the largest product of a list IS the own factor times
  // the largest opponent factor, so this is K1's per-list maximum, found without its second search pass)
  float best[3] = {0.0f, 0.0f, 0.0f};
  uint32_t bestk[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu};  // piece offset << 8 | entry (pieces are < 256 entries)
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
        const uint32_t ent = a;
        a += L.n0 * ESZ;
        if (L.flags & kTeamFirst) {
#pragma unroll
          for (int e = 0; e < 6; e++) g[e] = 0.0f;
        }
        float2 own = row(L.own_off);
        if (L.kind >= kH0) { const float s = own.x; own.x = own.y; own.y = s; }
        const bool home = (L.kind & 1) == 0;
        const float oxy = own.x * own.y;
        float gx, gy;
        if (!CLIP) {
          float ax = 0.0f, ay = 0.0f;
          for (uint32_t e = lane; e < L.n0; e += 32) {
            const uint2 q = lds64u(ent + e * 8);
            const float2 ea = row(q.x);
            const float w = __uint_as_float(q.y);
            ax = fmaf(w, ea.x, ax);
            ay = fmaf(w, ea.y, ay);
            if (home) {
              const float v0 = own.x * ea.x, v1 = own.y * ea.y, v2 = oxy * (ea.x * ea.y);
              const uint32_t key = (hoff << 8) | e;
              if (v0 > best[0]) { best[0] = v0; bestk[0] = key; }
              if (v1 > best[1]) { best[1] = v1; bestk[1] = key; }
              if (v2 > best[2]) { best[2] = v2; bestk[2] = key; }
            }
          }
          ax = wsum(ax);
          ay = wsum(ay);
          const float SX = own.x * ax, SY = own.y * ay;
          if (lane == 0) lp_acc -= 0.5f * (SX + SY);  // every match is in two lists
          gx = -SX;
          gy = -SY;
        } else {
          float lgx = 0.0f, lgy = 0.0f, lwx = 0.0f, lwy = 0.0f;
          gx = gy = 0.0f;
          for (uint32_t e = lane; e < L.n0; e += 32) {
            const uint4 q = lds128u(ent + e * 16);
            const float2 ea = row(q.x);
            const float w = __uint_as_float(q.y), wyx = __uint_as_float(q.z), wyy = __uint_as_float(q.w);
            const float X = own.x * ea.x, Y = own.y * ea.y;
            const float dx = fmaf(-w, X, wyx), dy = fmaf(-w, Y, wyy);  // w (y - rate): the gradient term of an unclipped rate
            if (X < 15.0f) gx += dx;
            if (Y < 15.0f) gy += dy;
            if (home) {
              const float cx = fminf(X, 15.0f), cy = fminf(Y, 15.0f), cp = cx * cy;
              lgx = fmaf(wyx, lg2_approx(cx), lgx);
              lgy = fmaf(wyy, lg2_approx(cy), lgy);
              lwx = fmaf(w, cx, lwx);
              lwy = fmaf(w, cy, lwy);
              const uint32_t key = (hoff << 8) | e;
              if (cx > best[0]) { best[0] = cx; bestk[0] = key; }
              if (cy > best[1]) { best[1] = cy; bestk[1] = key; }
              if (cp > best[2]) { best[2] = cp; bestk[2] = key; }
            }
          }
          gx = wsum(gx);
          gy = wsum(gy);
          if (home) {
            const float lg = wsum(lgx + lgy), lw = wsum(lwx + lwy);
            if (lane == 0) lp_acc += fmaf(lg, kLn2, -lw);
          }
        }
        add_own(g, L.kind, gx, gy);
        if (kp.Cf > 0) {
          cacc += L.kind == kH1 ? gx - gy : gy - gx;  // d/d (A - B) of the virtual team
          if (L.flags & kVteamLast) {
            if (lane == 0) S.conf1[L.vteam] = cacc;
            cacc = 0.0f;
          }
        }
        if (L.flags & kTeamLast) {
          if (dc) hacc += g[eAh1];
          if (lane == 0) {
            float* q = S.raw1 + (size_t)L.team * 8;
            q[0] = g[eAh1] + g[eAa1] + g[eA0];
            q[1] = -(g[eBh1] + g[eBa1] + g[eB0]);
            q[2] = g[eAh1];
            q[3] = g[eAa1];
            q[4] = -g[eBh1];
            q[5] = -g[eBa1];
          }
        }
      }
      ring.release(k);
    }
  }
  const uint32_t b2_0 = wb.b2[warp], b2_1 = wb.b2[warp + 1];
  ring.begin(kp.stream2 + b2_0, b2_1 - b2_0);  // tau pieces start streaming in during the bounds step
  // ---- bounds (bpl/_util.py:17-31): maxima over lanes and warps, lowest (piece, entry) among equals -----------------
#pragma unroll
  for (int q = 0; q < 3; q++) wargmax(best[q], bestk[q]);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 3; q++) {
      S.bestv[warp * 3 + q] = best[q];
      S.besti[warp * 3 + q] = bestk[q];
    }
  }
  __syncthreads();
  uint32_t besth[3], beste[3];
#pragma unroll
  for (int q = 0; q < 3; q++) {
    float bv = 0.0f;
    uint32_t bk = 0xffffffffu;
    for (int w = 0; w < W; w++) {
      const float v = S.bestv[w * 3 + q];
      const uint32_t kk = S.besti[w * 3 + q];
      if (v > bv || (v == bv && kk < bk)) {
        bv = v;
        bk = kk;
      }
    }
    best[q] = bv;
    besth[q] = bk >> 8;
    beste[q] = bk & 0xffu;
  }
  const float Lam = fmaxf(best[0], best[1]);
  const int qlam = best[0] >= best[1] ? 0 : 1;
  const float LB = -1.0f / Lam;
  const float UB = fminf(1.0f / best[2], 1.0f);
  const float cc = fmaf(r, UB - LB, LB);

  // ---- phase 2: tau pieces (bpl/_util.py:54-91), lanes over the entries ----------------------------------------
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
        uint32_t ent = a;
        a += (L.n0 + L.n1 + L.n2) * (uint32_t)sizeof(Entry);
        if (L.flags & kTeamFirst) {
#pragma unroll
          for (int e = 0; e < 6; e++) g[e] = 0.0f;
        }
        float2 own = row(L.own_off);
        if (L.kind >= kH0) { const float s = own.x; own.x = own.y; own.y = s; }
        const bool home = (L.kind & 1) == 0;
        float lt = 0.0f, uxy = 0.0f, sxx = 0.0f, sxy = 0.0f, u1x = 0.0f, u1y = 0.0f, s1x = 0.0f, s1y = 0.0f;
        for (uint32_t e = lane; e < L.n0; e += 32) {  // tau = 1 - c X Y
          const uint2 q = lds64u(ent + e * 8);
          const float2 ea = row(q.x);
          const float w = __uint_as_float(q.y);
          const float X = own.x * ea.x, Y = own.y * ea.y;  // the two rates first, then their product (overflow far out)
          const float t = CLIP ? fminf(X, 15.0f) * fminf(Y, 15.0f) : X * Y;
          const float tau = fmaxf(fmaf(-cc, t, 1.0f), 0.0f);
          const float wt = w * t, rc = rcp_approx(tau);
          uxy = fmaf(wt, rc, uxy);
          if (CLIP) {
            if (X < 15.0f) sxx = fmaf(wt, rc, sxx);
            if (Y < 15.0f) sxy = fmaf(wt, rc, sxy);
          }
          if (home) lt = fmaf(w, lg2_approx(tau), lt);
        }
        ent += L.n0 * 8;
#pragma unroll
        for (int c = 0; c < 2; c++) {  // tau = 1 + c X, then tau = 1 + c Y
          const uint32_t n = c == 0 ? L.n1 : L.n2;
          const float oc = c == 0 ? own.x : own.y;
          float uu = 0.0f, ss = 0.0f;
          for (uint32_t e = lane; e < n; e += 32) {
            const uint2 q = lds64u(ent + e * 8);
            const float2 ea = row(q.x);
            const float w = __uint_as_float(q.y);
            const float Rr = oc * ((q.x & 4u) ? ea.y : ea.x);  // `off` selects the .x or the .y float of the row
            const float R = CLIP ? fminf(Rr, 15.0f) : Rr;
            const float tau = fmaxf(fmaf(cc, R, 1.0f), 0.0f);
            const float wr = w * R, rc = rcp_approx(tau);
            uu = fmaf(wr, rc, uu);
            if (CLIP && Rr < 15.0f) ss = fmaf(wr, rc, ss);
            if (home) lt = fmaf(w, lg2_approx(tau), lt);
          }
          ent += n * 8;
          if (c == 0) { u1x = uu; s1x = ss; }
          else { u1y = uu; s1y = ss; }
        }
        uxy = wsum(uxy);
        u1x = wsum(u1x);
        u1y = wsum(u1y);
        float gx, gy;
        if (CLIP) {
          sxx = wsum(sxx); sxy = wsum(sxy); s1x = wsum(s1x); s1y = wsum(s1y);
          gx = cc * (s1x - sxx);
          gy = cc * (s1y - sxy);
        } else {
          gx = cc * (u1x - uxy);
          gy = cc * (u1y - uxy);
        }
        if (home) {
          lt = wsum(lt);
          if (lane == 0) lp_acc = fmaf(lt, kLn2, lp_acc);
          gc += u1x + u1y - uxy;  // (warp-uniform; lane 0's copy is the one that counts)
        }
        add_own(g, L.kind, gx, gy);
        if (kp.Cf > 0) {
          cacc += L.kind == kH1 ? gx - gy : gy - gx;
          if (L.flags & kVteamLast) {
            if (lane == 0) S.conf2[(L.kind & 1) * (V + 1) + L.vteam] += cacc;
            cacc = 0.0f;
          }
        }
        if (L.flags & kTeamLast) {
          if (dc) hacc += g[eAh1];
          if (lane == 0) {
            float* q = S.raw2 + ((size_t)(L.kind & 1) * T + L.team) * 8;
            q[0] = g[eAh1] + g[eAa1] + g[eA0];
            q[1] = -(g[eBh1] + g[eBa1] + g[eB0]);
            q[2] = g[eAh1];
            q[3] = g[eAa1];
            q[4] = -g[eBh1];
            q[5] = -g[eBa1];
          }
        }
      }
      ring.release(k);
    }
  }
  {  // d/d corr_coef and d/d home_advantage over the warps
    float v[2] = {lane == 0 ? gc : 0.0f, lane == 0 ? hacc : 0.0f};
    block_sum(v, 2);
    gc = S.tot[0];
    hacc = S.tot[1];
  }
  {  // the 1-1 matches: tau = 1 - c for all of them
    const float t11 = fmaxf(1.0f - cc, 0.0f);
    gc -= kp.w11 / t11;
    if (tid == 0 && kp.w11 != 0.0f) lp_acc = fmaf(kp.w11, logf(t11), lp_acc);
  }
  // ---- arg-max fix-up: the two matches, known to every thread -----------------------------
  Fixup fx;
  fx.h1 = 0u;
  fx.confs = 0u;
#pragma unroll
  for (int which = 0; which < 2; which++) {
    fx.teams[which] = 0xffffffffu;
    fx.vts[which] = 0xffffffffu;
    fx.vx[which] = fx.vy[which] = 0.0f;
    const int q = which == 0 ? qlam : 2;
    const bool need = (which == 0 || best[2] > 1.0f) && best[q] > 0.0f;
    if (need) {
      const unsigned char* base = kp.stream1 + besth[q];
      const Hdr L = unpack_hdr(__ldg(reinterpret_cast<const uint4*>(base)));
      const uint32_t off = __ldg(reinterpret_cast<const uint32_t*>(base + 16 + (size_t)beste[q] * ESZ));
      const bool h1 = L.kind == kH1;
      float2 own = row(L.own_off);
      if (L.kind >= kH0) { const float s = own.x; own.x = own.y; own.y = s; }
      const float2 ea = row(off);
      const bool xfree = !CLIP || own.x * ea.x < 15.0f, yfree = !CLIP || own.y * ea.y < 15.0f;
      const uint32_t own_v = L.vteam, opp_v = ((off >> 8) - ((h1 ? kp.tabQ1 : kp.tabP0) >> 8));
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
  // ---- team pass: one thread per team ---------------------------------------------------------------------------
  float rho = 0.0f, inv_s2 = 1.0f;
  if (has_rho) {
    rho = 2.0f * u - 1.0f;
    inv_s2 = 1.0f / (1.0f - rho * rho);
  }
  float acc[13];
#pragma unroll
  for (int i = 0; i < 13; i++) acc[i] = 0.0f;  // 0 lp | 1 mu_d | 2 ls_a | 3 ls_d | 4..7 mu | 8..11 ls | 12 rho
  for (int t = tid; t < T; t += NT) {
    const float za = ln.ld(o.za + t), zd = ln.ld(o.zd + t);
    const uint32_t v0 = (uint32_t)__ldg(kp.team_vptr + t), nv = (uint32_t)__ldg(kp.team_vptr + t + 1) - v0;
    const float* ys = kp.yteam + (size_t)t * 8;
    const float* q1 = S.raw1 + (size_t)t * 8;
    const float* q2 = S.raw2 + (size_t)t * 8;
    const float* q3 = S.raw2 + ((size_t)T + t) * 8;
    float ra = q1[0] + __ldg(ys) + q2[0] + q3[0];
    float rd = q1[1] + __ldg(ys + 1) + q2[1] + q3[1];
    float rx[4], dec[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int i = 0; i < 4; i++) rx[i] = __ldg(ys + 2 + i);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (i < ndec) {
        dec[i] = ln.ld(o.dec[i] + t);
        rx[i] += q1[2 + i] + q2[2 + i] + q3[2 + i];
      }
    }
#pragma unroll
    for (int which = 0; which < 2; which++) {
      if ((fx.vts[which] & 0xffffu) - v0 < nv) fold_fixup(fx, which, 0, ra, rd, rx);
      if ((fx.vts[which] >> 16) - v0 < nv) fold_fixup(fx, which, 1, ra, rd, rx);
    }
    float p_za, p_zd;
    if (has_rho) {  // za ~ N(0,1), zd ~ N(rho za, sqrt(1-rho^2))  (extended_dixon_coles.py:165-174)
      const float e = zd - rho * za;
      const float es = e * inv_s2;
      acc[0] -= 0.5f * (za * za + e * es);
      p_za = -za + rho * es;
      p_zd = -es;
      acc[12] += es * za - rho * es * es + rho * inv_s2;
    } else {
      acc[0] -= pri * 0.5f * (za * za + zd * zd);
      p_za = -pri * za;
      p_zd = -pri * zd;
    }
    *ln.g(o.za + t) = fmaf(hy.sig_a, ra, p_za);
    *ln.g(o.zd + t) = fmaf(hy.sig_d, rd, p_zd);
    acc[2] = fmaf(hy.sig_a * za, ra, acc[2]);
    acc[3] = fmaf(hy.sig_d * zd, rd, acc[3]);
    acc[1] += rd;
    if (dc) acc[4] += rx[0];  // (only the fix-up's share and the host-folded statics: the lists' share is hacc)
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (i < ndec) {
        acc[0] -= pri * 0.5f * dec[i] * dec[i];
        *ln.g(o.dec[i] + t) = fmaf(hy.sig[i], rx[i], -pri * dec[i]);
        acc[4 + i] += rx[i];
        acc[8 + i] = fmaf(hy.sig[i] * dec[i], rx[i], acc[8 + i]);
      }
    }
    if (kp.K > 0) {
      S.rows[t * 2] = ra;
      S.rows[t * 2 + 1] = rd;
    }
  }
  // confederation strengths: N(0,1) prior + sum over the virtual teams of the confederation
  for (int k = tid; k < kp.Cf; k += NT) {
    const float cf = ln.ld(o.conf + k);
    float s = __ldg(kp.yconf + k) - pri * cf;
    acc[0] -= pri * 0.5f * cf * cf;
    const int j0 = __ldg(kp.conf_vptr + k), j1 = __ldg(kp.conf_vptr + k + 1);
    for (int j = j0; j < j1; j++) {
      const int v = __ldg(kp.conf_vlist + j);
      s += S.conf1[v] + S.conf2[v] + S.conf2[(V + 1) + v];
    }
#pragma unroll
    for (int ws = 0; ws < 4; ws++)
      if (fx.vts[ws >> 1] != 0xffffffffu && ((fx.confs >> (8 * ((ws >> 1) * 2 + (ws & 1)))) & 0xffu) == (uint32_t)k)
        s += fixup_conf(fx, ws >> 1, ws & 1);
    *ln.g(o.conf + k) = s;
  }
  acc[0] += lp_acc;
  block_sum(acc, 13);
  // ---- epilogue: the scalar sites, one thread each ----------------------------------------------------------------
  float lp = 0.0f;
  const int n_items = kp.nhyper + 2 + 2 * kp.K;
  for (int item = tid; item < n_items; item += NT) {
    if (item < kp.nhyper) {
      const HyperDesc hd = kp.hyper[item];
      const float x = ln.ld(hd.off);
      const float a = S.tot[hd.row] + ((dc && hd.row == 4) ? hacc : 0.0f);
      float gval;
      if (hd.kind == 0) {  // Normal(loc, scale)
        const float z = (x - hd.loc) * hd.inv_scale;
        lp -= 0.5f * z * z;
        gval = fmaf(-z, hd.inv_scale, a);
      } else if (hd.kind == 2) {  // likelihood-only: no prior
        gval = a;
      } else {  // HalfNormal(scale) on exp(x) + Jacobian x
        const float z = expf(x) * hd.inv_scale;
        lp += fmaf(-0.5f * z, z, x);
        gval = fmaf(-z, z, 1.0f) + a;
      }
      *ln.g(hd.off) = gval;
    } else if (item == kp.nhyper) {
      if (has_rho) {  // u ~ Beta(2,4) + sigmoid Jacobian; rho = 2u - 1; sum_t -log sqrt(1 - rho^2)
        lp += 0.5f * (float)T * logf(inv_s2) + 2.0f * logf(u) + 4.0f * logf(1.0f - u);
        *ln.g(o.u) = 2.0f - 6.0f * u + S.tot[12] * 2.0f * u * (1.0f - u);
      }
    } else if (item == kp.nhyper + 1) {
      // corr_coef_raw ~ Beta(2,2) + Jacobian; corr_coef = LB + r (UB - LB); also the likelihood + team-prior sum
      lp += S.tot[0] + (lik ? 0.0f : 2.0f * (logf(r) + logf(1.0f - r)));
      *ln.g(o.raw) = lik ? gc * (UB - LB) : 2.0f * (1.0f - 2.0f * r) + gc * r * (1.0f - r) * (UB - LB);
      if (kp.corr_coef) kp.corr_coef[chain] = cc;
    } else {  // covariate coefficients: d/d beta[k] = sum_t Xs[t,k] * d/d (att | def)[t]; N(0,1) prior
      const int task = item - kp.nhyper - 2, k = task >> 1, isd = task & 1;
      const int d = (isd ? o.beta_d : o.beta_a) + k;
      const float b = ln.ld(d);
      float s = -b;
      lp -= 0.5f * b * b;
      for (int t = 0; t < T; t++) s = fmaf(__ldg(kp.Xs + (size_t)t * kp.K + k), S.rows[t * 2 + isd], s);
      *ln.g(d) = s;
    }
  }
  {
    float v[1] = {lp};
    block_sum(v, 1);
    if (tid == 0) {
      lp = kp.const_term + S.tot[0];
      kp.lp[chain] = lp == INFINITY ? NAN : lp;  // (see logdensity.cu: +inf is an overflowed intermediate)
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------
static size_t few_smem_bytes(const KernelParams& kp) {
  const size_t nrows = (kp.tab_bytes >> 8) + 1;
  return (size_t)kp.nwarps * kStages * (kp.stage_bytes + 8) + nrows * 8 + (size_t)kp.T * (32 + 64 + 8) +
         (size_t)(kp.V + 1) * 12 + (size_t)kp.nwarps * (64 + 12 + 12) + 64 + 64 + 64;
}

bool logdensity_few_supported(const KernelParams& kp) {
  return kp.model != BPLX_DYNAMIC && few_smem_bytes(kp) <= 200 * 1024;
}

int launch_logdensity_few(const KernelParams& kp, const WarpBounds& wb, cudaStream_t stream) {
  const size_t smem = few_smem_bytes(kp);
  static bool attr_done[2] = {false, false};
  if (!attr_done[kp.clip ? 1 : 0] && smem > 48 * 1024) {
    if (kp.clip) BPLX_CUDA(cudaFuncSetAttribute(&logdensity_few_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    else BPLX_CUDA(cudaFuncSetAttribute(&logdensity_few_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done[kp.clip ? 1 : 0] = true;
  }
  if (kp.clip) logdensity_few_kernel<true><<<kp.C, kp.nwarps * 32, smem, stream>>>(kp, wb);
  else logdensity_few_kernel<false><<<kp.C, kp.nwarps * 32, smem, stream>>>(kp, wb);
  BPLX_CUDA(cudaGetLastError());
  note_launch(1);
  return BPLX_OK;
}

}  // namespace bplx
