// logdensity.cu -- K1: fused Dixon-Coles-family log-density + gradient, one lane per chain.
//
// Replaces value_and_grad(potential_fn) of the reference `_model`s (bpl/dixon_coles.py:39-84,
// bpl/extended_dixon_coles.py:78-248, bpl/neutral_dixon_coles.py:102-283,
// bpl/neutral_dixon_coles_WC.py:83-232) together with bpl/_util.py:17-93, for a batch of chains.
//
// CTA = 32 chains (lane = chain) x nwarps warps.  Flow (DESIGN.md "K1"):
//   prologue  each warp takes teams t = warp, warp+W, ...: reads theta, adds the per-team priors,
//             writes the rows (exp of the six exponents) of its virtual teams into the shared
//             tables, and stores the prior + static (sum w*y) part of the gradient.
//   phase 1   each warp walks its entry lists: acc += w * table[opp]; per list
//             sum_w_lambda = own * acc (both rates), running maxima of lambda_h, lambda_a,
//             lambda_h*lambda_a with the list they came from (bpl/_util.py:23-30).
//   bounds    maxima reduced over warps -> LB, UB, corr_coef.
//   phase 2   tau lists (0-0, 1-0, 0-1 matches): log tau, d/d eta, d/d corr_coef.
//   fix-up    the maxima's arg-max entries are found by rescanning one list per chain and the
//             d corr_coef / d eta terms are added.
//   epilogue  hyper-parameter chain rule, priors and Jacobians; lp, corr_coef.
// Gradients of per-team parameters are accumulated by read-modify-write of the caller's grad
// buffer by the single thread that owns (team, chain) in each stage: no atomics, deterministic.
#include <float.h>

#include "common.cuh"
#include "plan.h"
#include "problem.h"

namespace bplx {

namespace {

constexpr float kLogSqrt2Pi = 0.918938533204672742f;
constexpr float kLn2 = 0.693147180559945309f;

struct Hyp {  // constrained hyper-parameters of this lane's chain
  float mu_d, sig_a, sig_d, mu[4], sig[4], rho, inv_s2;
};

template <int KMAX>
struct Acc {  // gradient accumulators wrt hyper-parameters (+ this thread's share of lp)
  float lp, mu_d, ls_a, ls_d, mu[4], ls[4], rho;
  float ba[KMAX > 0 ? KMAX : 1], bd[KMAX > 0 ? KMAX : 1];
};

struct Lane {
  const float* th;  // theta + chain * sc
  float* gr;        // grad  + chain * sc
  float* sc;        // scratch + chain
  long long sd;
  bool active;
  __device__ __forceinline__ float ld(int d) const { return __ldg(th + (long long)d * sd); }
};

__device__ __forceinline__ float sigmoid_clipped(float x) {
  // numpyro SigmoidTransform: clip(expit(x), finfo.tiny, 1 - finfo.eps)
  float s = 1.0f / (1.0f + expf(-x));
  return fminf(fmaxf(s, FLT_MIN), 1.0f - FLT_EPSILON);
}

// exponent gradients g[6] of virtual team v -> raw parameter gradients.
// `pred` masks everything (stores and accumulators); `init_t` / `init_v`: first touch of the
// team's gradient entries / of the virtual team's scratch slot -> plain store of prior + value.
template <int KMAX>
__device__ __forceinline__ void apply_vteam(const KernelParams& kp, const Hyp& hy, Acc<KMAX>& acc, const Lane& ln,
                                            bool pred, int v, const float (&g)[6], bool init_t, bool init_v,
                                            float p_za, float p_zd, const float (&p_dec)[4]) {
  const ThetaOffsets& o = kp.off;
  const int t = kp.v_team[v];
  const float gA = g[eAh1] + g[eAa1] + g[eA0];
  const float gB = g[eBh1] + g[eBa1] + g[eB0];
  const float g_att = pred ? gA : 0.0f, g_def = pred ? -gB : 0.0f;
  const bool st = pred && ln.active;
  if (kp.Cf > 0) {
    float* s = ln.sc + (long long)v * kp.Cpad;
    float old = init_v ? 0.0f : (st ? *s : 0.0f);
    if (st) *s = old + (gA - gB);
  }
  {
    const float za = ln.ld(o.za + t), zd = ln.ld(o.zd + t);
    float* pa = ln.gr + (long long)(o.za + t) * ln.sd;
    float* pd = ln.gr + (long long)(o.zd + t) * ln.sd;
    float olda = init_t ? p_za : (st ? *pa : 0.0f);
    float oldd = init_t ? p_zd : (st ? *pd : 0.0f);
    if (st) {
      *pa = fmaf(hy.sig_a, g_att, olda);
      *pd = fmaf(hy.sig_d, g_def, oldd);
    }
    acc.ls_a = fmaf(hy.sig_a * za, g_att, acc.ls_a);
    acc.ls_d = fmaf(hy.sig_d * zd, g_def, acc.ls_d);
    acc.mu_d += g_def;
  }
  if (KMAX > 0) {
#pragma unroll
    for (int k = 0; k < KMAX; k++) {
      if (k < kp.K) {
        const float x = __ldg(kp.Xs + (size_t)t * kp.K + k);
        acc.ba[k] = fmaf(x, g_att, acc.ba[k]);
        acc.bd[k] = fmaf(x, g_def, acc.bd[k]);
      }
    }
  }
  const float gx[4] = {pred ? g[eAh1] : 0.0f, pred ? g[eAa1] : 0.0f, pred ? -g[eBh1] : 0.0f, pred ? -g[eBa1] : 0.0f};
  if (kp.model == BPLX_DIXON_COLES) {
    acc.mu[0] += gx[0];
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (i == 0 || kp.model != BPLX_EXTENDED) {
        const float dec = ln.ld(o.dec[i] + t);
        float* p = ln.gr + (long long)(o.dec[i] + t) * ln.sd;
        float old = init_t ? p_dec[i] : (st ? *p : 0.0f);
        if (st) *p = fmaf(hy.sig[i], gx[i], old);
        acc.mu[i] += gx[i];
        acc.ls[i] = fmaf(hy.sig[i] * dec, gx[i], acc.ls[i]);
      }
    }
  }
}

__device__ __forceinline__ void add_own(float (&g)[6], int kind, float gx, float gy) {
  switch (kind) {
    case kH1: g[eAh1] += gx; g[eBh1] += gy; break;
    case kA1: g[eBa1] += gx; g[eAa1] += gy; break;
    default: g[eB0] += gx; g[eA0] += gy; break;
  }
}

__device__ __forceinline__ List load_list(const List* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  List L;
  L.ent = a.x; L.n = a.y; L.own_off = a.z; L.vteam = a.w;
  L.n_xy = (uint16_t)(b.x & 0xffff); L.n_x = (uint16_t)(b.x >> 16);
  L.n_y = (uint16_t)(b.y & 0xffff); L.kind = (uint8_t)((b.y >> 16) & 0xff); L.flags = (uint8_t)(b.y >> 24);
  return L;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Per-warp TMA ring over a contiguous global byte stream (the warp's entries of one phase).
// One elected lane issues cp.async.bulk copies of kStageBytes into the warp's private ring and
// every lane waits on the stage's mbarrier before reading it; the same warp produces and
// consumes, so a __syncwarp() is all that is needed before a slot is refilled.
struct Stream {
  uint32_t ring, bar;        // shared addresses: kStages * kStageBytes ring, kStages mbarriers
  const unsigned char* src;  // current stream
  uint32_t total, pos;       // bytes in / consumed from the current stream
  uint32_t gs0, gs_next;     // ring-stage counter at the start of / after the current stream
  int lane;

  __device__ __forceinline__ void init(uint32_t ring_, uint32_t bar_, int lane_) {
    ring = ring_; bar = bar_; lane = lane_;
    total = pos = gs0 = gs_next = 0;
    src = nullptr;
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < kStages; s++) mbar_init(bar + 8 * s, 1);
      fence_mbar_init();
      fence_proxy_async();
    }
    __syncwarp();
  }
  __device__ __forceinline__ void issue(uint32_t k) {  // stage k of the current stream
    const uint32_t b0 = k * kStageBytes;
    if (lane == 0 && b0 < total) {
      const uint32_t bytes = min((uint32_t)kStageBytes, total - b0);
      const uint32_t slot = (gs0 + k) % kStages;
      mbar_arrive_expect_tx(bar + 8 * slot, bytes);
      tma_load_1d(ring + slot * kStageBytes, src + b0, bytes, bar + 8 * slot);
    }
  }
  __device__ __forceinline__ void begin(const void* src_, uint32_t total_bytes) {
    __syncwarp();  // every lane is done with the previous stream's stages
    src = static_cast<const unsigned char*>(src_);
    total = total_bytes;
    pos = 0;
    gs0 = gs_next;
    gs_next = gs0 + (total_bytes + kStageBytes - 1) / kStageBytes;
#pragma unroll
    for (int s = 0; s < kStages; s++) issue(s);
  }
  // body(shared address of a 16-byte unit), for nbytes (multiple of 16) of the stream
  template <typename F>
  __device__ __forceinline__ void consume(uint32_t nbytes, F&& body) {
    while (nbytes) {
      const uint32_t in_stage = pos & (kStageBytes - 1);
      const uint32_t k = pos / kStageBytes;
      const uint32_t idx = gs0 + k;
      const uint32_t slot = idx % kStages;
      if (in_stage == 0) mbar_wait(bar + 8 * slot, (idx / kStages) & 1);
      const uint32_t chunk = min(nbytes, (uint32_t)kStageBytes - in_stage);
      const uint32_t a0 = ring + slot * kStageBytes + in_stage;
#pragma unroll 4
      for (uint32_t o = 0; o < chunk; o += 16) body(a0 + o);
      pos += chunk;
      nbytes -= chunk;
      if ((pos & (kStageBytes - 1)) == 0) {
        __syncwarp();
        issue(k + kStages);
      }
    }
  }
};

}  // namespace

template <bool CLIP, int KMAX>
__global__ void __launch_bounds__(kMaxWarps * 32, 1) logdensity_kernel(const __grid_constant__ KernelParams kp) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = kp.nwarps;
  const ThetaOffsets& o = kp.off;
  const int chain_raw = blockIdx.x * kChains + lane;
  const int chain = min(chain_raw, kp.C - 1);
  Lane ln;
  ln.th = kp.theta + (long long)chain * kp.sc;
  ln.gr = kp.grad + (long long)chain * kp.sc;
  ln.sc = kp.scratch + chain;
  ln.sd = kp.sd;
  ln.active = chain_raw < kp.C;
  const uint32_t tab = smem_u32(smem) + lane * 8;  // + row byte offset
  float* red = reinterpret_cast<float*>(smem + kp.smem_red);
  Stream stream;
  stream.init(smem_u32(smem) + kp.smem_ring + warp * (kStages * kStageBytes),
              smem_u32(smem) + kp.smem_bar + warp * (kStages * 8), lane);
  {  // phase-1 entries start streaming in while the prologue runs
    const int e0 = __ldg(kp.warp_e1 + warp), e1 = __ldg(kp.warp_e1 + warp + 1);
    const uint32_t esz = CLIP ? (uint32_t)sizeof(EntryClip) : (uint32_t)sizeof(Entry);
    stream.begin(static_cast<const unsigned char*>(kp.ent1) + (size_t)e0 * esz, (uint32_t)(e1 - e0) * esz);
  }

  const bool dc = kp.model == BPLX_DIXON_COLES, ext = kp.model == BPLX_EXTENDED;
  const bool has_rho = !dc;
  const int ndec = dc ? 0 : (ext ? 1 : 4);

  // ---- hyper-parameters ---------------------------------------------------------------------
  Hyp hy;
  hy.mu_d = ln.ld(o.mean_defence);
  hy.sig_a = expf(ln.ld(o.log_std_attack));
  hy.sig_d = expf(ln.ld(o.log_std_defence));
#pragma unroll
  for (int i = 0; i < 4; i++) {
    hy.mu[i] = o.mean[i] >= 0 ? ln.ld(o.mean[i]) : 0.0f;
    hy.sig[i] = o.log_std[i] >= 0 ? expf(ln.ld(o.log_std[i])) : 0.0f;
  }
  float u = 0.5f;
  if (has_rho) u = sigmoid_clipped(ln.ld(o.u));
  hy.rho = has_rho ? 2.0f * u - 1.0f : 0.0f;
  hy.inv_s2 = 1.0f / (1.0f - hy.rho * hy.rho);
  float beta_a[KMAX > 0 ? KMAX : 1], beta_d[KMAX > 0 ? KMAX : 1];
  if (KMAX > 0) {
#pragma unroll
    for (int k = 0; k < KMAX; k++) {
      beta_a[k] = k < kp.K ? ln.ld(o.beta_a + k) : 0.0f;
      beta_d[k] = k < kp.K ? ln.ld(o.beta_d + k) : 0.0f;
    }
  }
  Acc<KMAX> acc;
  acc.lp = acc.mu_d = acc.ls_a = acc.ls_d = acc.rho = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; i++) acc.mu[i] = acc.ls[i] = 0.0f;
#pragma unroll
  for (int k = 0; k < (KMAX > 0 ? KMAX : 1); k++) acc.ba[k] = acc.bd[k] = 0.0f;

  // ---- prologue -------------------------------------------------------------------------------
  if (warp == 0) {  // zero rows used by padding entries
    const uint32_t zr = (uint32_t)kp.V * kRowBytes;
    if (kp.has1) {
      sts64(tab + kp.tabP1 + zr, 0.0f, 0.0f);
      sts64(tab + kp.tabQ1 + zr, 0.0f, 0.0f);
    }
    if (kp.has0) sts64(tab + kp.tabP0 + zr, 0.0f, 0.0f);
  }
  for (int t = warp; t < kp.T; t += W) {
    const float za = ln.ld(o.za + t), zd = ln.ld(o.zd + t);
    float am = 0.0f, dm = hy.mu_d;
    if (KMAX > 0) {
#pragma unroll
      for (int k = 0; k < KMAX; k++) {
        if (k < kp.K) {
          const float x = __ldg(kp.Xs + (size_t)t * kp.K + k);
          am = fmaf(x, beta_a[k], am);
          dm = fmaf(x, beta_d[k], dm);
        }
      }
    }
    const float att = fmaf(za, hy.sig_a, am), def = fmaf(zd, hy.sig_d, dm);
    float p_za, p_zd, p_dec[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (has_rho) {  // za ~ N(0,1), zd ~ N(rho za, sqrt(1-rho^2))  (extended_dixon_coles.py:165-174)
      const float e = zd - hy.rho * za;
      const float es = e * hy.inv_s2;
      acc.lp -= 0.5f * (za * za + e * es);
      p_za = -za + hy.rho * es;
      p_zd = -es;
      acc.rho += es * za - hy.rho * es * es + hy.rho * hy.inv_s2;
    } else {
      acc.lp -= 0.5f * (za * za + zd * zd);
      p_za = -za;
      p_zd = -zd;
    }
    float x[4] = {dc ? hy.mu[0] : 0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (i < ndec) {
        const float dec = ln.ld(o.dec[i] + t);
        x[i] = fmaf(hy.sig[i], dec, hy.mu[i]);
        acc.lp -= 0.5f * dec * dec;
        p_dec[i] = -dec;
      }
    }
    const int v0 = __ldg(kp.team_vptr + t), v1 = __ldg(kp.team_vptr + t + 1);
    if (v0 == v1 && ln.active) {  // team without matches: prior gradient only
      ln.gr[(long long)(o.za + t) * ln.sd] = p_za;
      ln.gr[(long long)(o.zd + t) * ln.sd] = p_zd;
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (i < ndec) ln.gr[(long long)(o.dec[i] + t) * ln.sd] = p_dec[i];
    }
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
      float g[6];
#pragma unroll
      for (int e = 0; e < 6; e++) {
        g[e] = CLIP ? 0.0f : __ldg(kp.yexp + (size_t)v * 6 + e);
        acc.lp = fmaf(g[e], ex[e], acc.lp);
      }
      apply_vteam<KMAX>(kp, hy, acc, ln, true, v, g, v == v0, true, p_za, p_zd, p_dec);
    }
  }
  __syncthreads();

  // ---- phase 1 ----------------------------------------------------------------------------------
  float best[3] = {0.0f, 0.0f, 0.0f};
  int bestl[3] = {0, 0, 0};
  {
    float g[6];
    const float zero4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const int l0 = __ldg(kp.warp_l1 + warp), l1 = __ldg(kp.warp_l1 + warp + 1);
    List Lnext = l0 < l1 ? load_list(kp.lists1 + l0) : List{};
    for (int li = l0; li < l1; li++) {
      const List L = Lnext;
      if (li + 1 < l1) Lnext = load_list(kp.lists1 + li + 1);  // hide the header's latency behind this list
      if (L.flags & kListFirst) {
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
          stream.consume(L.n * (uint32_t)sizeof(Entry), [&](uint32_t addr) {
            const uint4 q = lds128u(addr);  // two entries
            const float2 a = lds64(tab + q.x), b = lds64(tab + q.z);
            const float wa = __uint_as_float(q.y), wb = __uint_as_float(q.w);
            ax0 = fmaf(wa, a.x, ax0); ay0 = fmaf(wa, a.y, ay0);
            ax1 = fmaf(wb, b.x, ax1); ay1 = fmaf(wb, b.y, ay1);
            m1 = fmaxf(m1, fmaxf(a.x, b.x));
            m2 = fmaxf(m2, fmaxf(a.y, b.y));
            m3 = fmaxf(m3, fmaxf(a.x * a.y, b.x * b.y));
          });
          const float v0 = own.x * m1, v1 = own.y * m2, v2 = (own.x * own.y) * m3;
          if (v0 > best[0]) { best[0] = v0; bestl[0] = li; }
          if (v1 > best[1]) { best[1] = v1; bestl[1] = li; }
          if (v2 > best[2]) { best[2] = v2; bestl[2] = li; }
        } else {
          stream.consume(L.n * (uint32_t)sizeof(Entry), [&](uint32_t addr) {
            const uint4 q = lds128u(addr);
            const float2 a = lds64(tab + q.x), b = lds64(tab + q.z);
            const float wa = __uint_as_float(q.y), wb = __uint_as_float(q.w);
            ax0 = fmaf(wa, a.x, ax0); ay0 = fmaf(wa, a.y, ay0);
            ax1 = fmaf(wb, b.x, ax1); ay1 = fmaf(wb, b.y, ay1);
          });
        }
        const float SX = own.x * (ax0 + ax1), SY = own.y * (ay0 + ay1);
        acc.lp -= 0.5f * (SX + SY);  // every match is in two lists
        gx = -SX;
        gy = -SY;
      } else {
        gx = gy = 0.0f;
        if (home) {
          float lp2 = 0.0f, lpw = 0.0f, m1 = 0.0f, m2 = 0.0f, m3 = 0.0f;
          stream.consume(L.n * (uint32_t)sizeof(EntryClip), [&](uint32_t addr) {
            const uint4 q = lds128u(addr);  // one entry: off, w, w*y_x, w*y_y
            const float2 a = lds64(tab + q.x);
            const float w = __uint_as_float(q.y), wyx = __uint_as_float(q.z), wyy = __uint_as_float(q.w);
            const float X = own.x * a.x, Y = own.y * a.y;
            const float Xc = fminf(X, 15.0f), Yc = fminf(Y, 15.0f);
            lp2 = fmaf(wyx, lg2_approx(Xc), lp2);
            lp2 = fmaf(wyy, lg2_approx(Yc), lp2);
            lpw = fmaf(w, Xc + Yc, lpw);
            m1 = fmaxf(m1, Xc);
            m2 = fmaxf(m2, Yc);
            m3 = fmaxf(m3, Xc * Yc);
            gx += X < 15.0f ? fmaf(-w, X, wyx) : 0.0f;
            gy += Y < 15.0f ? fmaf(-w, Y, wyy) : 0.0f;
          });
          acc.lp += fmaf(lp2, kLn2, -lpw);
          if (m1 > best[0]) { best[0] = m1; bestl[0] = li; }
          if (m2 > best[1]) { best[1] = m2; bestl[1] = li; }
          if (m3 > best[2]) { best[2] = m3; bestl[2] = li; }
        } else {
          stream.consume(L.n * (uint32_t)sizeof(EntryClip), [&](uint32_t addr) {
            const uint4 q = lds128u(addr);
            const float2 a = lds64(tab + q.x);
            const float w = __uint_as_float(q.y), wyx = __uint_as_float(q.z), wyy = __uint_as_float(q.w);
            const float X = own.x * a.x, Y = own.y * a.y;
            gx += X < 15.0f ? fmaf(-w, X, wyx) : 0.0f;
            gy += Y < 15.0f ? fmaf(-w, Y, wyy) : 0.0f;
          });
        }
      }
      add_own(g, L.kind, gx, gy);
      if (L.flags & kListLast) apply_vteam<KMAX>(kp, hy, acc, ln, true, (int)L.vteam, g, false, false, 0.0f, 0.0f, zero4);
    }
  }
  {  // start fetching this warp's tau entries while the other warps finish phase 1
    const int e0 = __ldg(kp.warp_e2 + warp), e1 = __ldg(kp.warp_e2 + warp + 1);
    stream.begin(kp.ent2 + e0, (uint32_t)(e1 - e0) * (uint32_t)sizeof(Entry));
  }
  // ---- bounds (bpl/_util.py:17-31) ------------------------------------------------------------------
#pragma unroll
  for (int q = 0; q < 3; q++) {
    red[(warp * kRedRows + q) * 32 + lane] = best[q];
    red[(warp * kRedRows + 3 + q) * 32 + lane] = __int_as_float(bestl[q]);
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 3; q++) {
    best[q] = 0.0f;
    bestl[q] = 0;
  }
  for (int w = 0; w < W; w++) {
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const float v = red[(w * kRedRows + q) * 32 + lane];
      const int l = __float_as_int(red[(w * kRedRows + 3 + q) * 32 + lane]);
      if (v > best[q]) { best[q] = v; bestl[q] = l; }
    }
  }
  const float Lam = fmaxf(best[0], best[1]);
  const int qlam = best[0] >= best[1] ? 0 : 1;
  const float LB = -1.0f / Lam;
  const float UB = fminf(1.0f / best[2], 1.0f);
  const float r = sigmoid_clipped(ln.ld(o.raw));
  const float cc = fmaf(r, UB - LB, LB);

  // ---- phase 2: tau terms (bpl/_util.py:54-91) -----------------------------------------------------
  float gc = 0.0f;
  {
    float g[6];
    const float zero4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const int l0 = __ldg(kp.warp_l2 + warp), l1 = __ldg(kp.warp_l2 + warp + 1);
    List Lnext = l0 < l1 ? load_list(kp.lists2 + l0) : List{};
    for (int li = l0; li < l1; li++) {
      const List L = Lnext;
      if (li + 1 < l1) Lnext = load_list(kp.lists2 + li + 1);
      if (L.flags & kListFirst) {
#pragma unroll
        for (int e = 0; e < 6; e++) g[e] = 0.0f;
      }
      float2 own = lds64(tab + L.own_off);
      if (L.kind >= kH0) { const float s = own.x; own.x = own.y; own.y = s; }
      const bool home = (L.kind & 1) == 0;
      float uxy = 0.0f, ux = 0.0f, uy = 0.0f, lt = 0.0f;     // unmasked (d/d corr_coef)
      float sxy_x = 0.0f, sxy_y = 0.0f, sx = 0.0f, sy = 0.0f;  // masked by "rate not clipped"
      // one entry = (opponent row offset, w); the ring is read two entries (16 bytes) at a time
      auto xy = [&](uint32_t off, float w) {  // tau = 1 - c X Y
        const float2 a = lds64(tab + off);
        const float Xr = own.x * a.x, Yr = own.y * a.y;
        const float X = CLIP ? fminf(Xr, 15.0f) : Xr, Y = CLIP ? fminf(Yr, 15.0f) : Yr;
        const float t = X * Y;
        const float tau = fmaxf(fmaf(-cc, t, 1.0f), 0.0f);
        const float val = (w * t) * rcp_approx(tau);
        uxy += val;
        if (CLIP) {
          sxy_x += Xr < 15.0f ? val : 0.0f;
          sxy_y += Yr < 15.0f ? val : 0.0f;
        }
        if (home) lt = fmaf(w, lg2_approx(tau), lt);
      };
      auto one = [&](uint32_t off, float w, bool is_x, float& u, float& s) {  // tau = 1 + c X  (or Y)
        const float2 a = lds64(tab + off);
        const float Rr = is_x ? own.x * a.x : own.y * a.y;
        const float R = CLIP ? fminf(Rr, 15.0f) : Rr;
        const float tau = fmaxf(fmaf(cc, R, 1.0f), 0.0f);
        const float val = (w * R) * rcp_approx(tau);
        u += val;
        if (CLIP) s += Rr < 15.0f ? val : 0.0f;
        if (home) lt = fmaf(w, lg2_approx(tau), lt);
      };
      stream.consume((uint32_t)L.n_xy * (uint32_t)sizeof(Entry), [&](uint32_t addr) {
        const uint4 q = lds128u(addr);
        xy(q.x, __uint_as_float(q.y));
        xy(q.z, __uint_as_float(q.w));
      });
      stream.consume((uint32_t)L.n_x * (uint32_t)sizeof(Entry), [&](uint32_t addr) {
        const uint4 q = lds128u(addr);
        one(q.x, __uint_as_float(q.y), true, ux, sx);
        one(q.z, __uint_as_float(q.w), true, ux, sx);
      });
      stream.consume((uint32_t)L.n_y * (uint32_t)sizeof(Entry), [&](uint32_t addr) {
        const uint4 q = lds128u(addr);
        one(q.x, __uint_as_float(q.y), false, uy, sy);
        one(q.z, __uint_as_float(q.w), false, uy, sy);
      });
      if (!CLIP) { sxy_x = sxy_y = uxy; sx = ux; sy = uy; }
      if (home) {
        acc.lp = fmaf(lt, kLn2, acc.lp);
        gc += ux + uy - uxy;
      }
      add_own(g, L.kind, cc * (sx - sxy_x), cc * (sy - sxy_y));
      if (L.flags & kListLast) apply_vteam<KMAX>(kp, hy, acc, ln, true, (int)L.vteam, g, false, false, 0.0f, 0.0f, zero4);
    }
  }
  red[(warp * kRedRows + 6) * 32 + lane] = gc;
  __syncthreads();
  gc = 0.0f;
  for (int w = 0; w < W; w++) gc += red[(w * kRedRows + 6) * 32 + lane];
  {  // the 1-1 matches: tau = 1 - c for all of them
    const float t11 = fmaxf(1.0f - cc, 0.0f);
    gc -= kp.w11 / t11;
    if (warp == 0 && kp.w11 != 0.0f) acc.lp = fmaf(kp.w11, logf(t11), acc.lp);
  }

  // ---- arg-max fix-up: chains l = warp, warp+W, ... ----------------------------
  {
    const int my_l[2] = {qlam == 0 ? bestl[0] : bestl[1], bestl[2]};
    const float my_t[2] = {Lam, best[2]};
    const float zero4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int l = warp; l < kChains; l += W) {
      const uint32_t tab_l = smem_u32(smem) + l * 8;
#pragma unroll 1
      for (int which = 0; which < 2; which++) {
        const int li = __shfl_sync(0xffffffffu, my_l[which], l);
        const float target = __shfl_sync(0xffffffffu, my_t[which], l);
        const int q = which == 0 ? __shfl_sync(0xffffffffu, qlam, l) : 2;
        if (which == 1 && !(target > 1.0f)) continue;  // UB = 1: no dependence on the rates
        const List L = load_list(kp.lists1 + li);
        float2 own = lds64(tab_l + L.own_off);
        if (L.kind >= kH0) { const float s = own.x; own.x = own.y; own.y = s; }
        uint32_t f_off = 0;
        float f_x = 0.0f, f_y = 0.0f;
        bool found = false;
        for (uint32_t i0 = 0; i0 < L.n && !found; i0 += 32) {
          const uint32_t i = i0 + lane;
          uint32_t off = 0;
          float X = 0.0f, Y = 0.0f, val = -1.0f;
          if (i < L.n) {
            off = CLIP ? __ldg(&reinterpret_cast<const EntryClip*>(kp.ent1)[L.ent + i].off)
                       : __ldg(&reinterpret_cast<const Entry*>(kp.ent1)[L.ent + i].off);
            const float2 a = lds64(tab_l + off);
            X = own.x * a.x;
            Y = own.y * a.y;
            if (CLIP) {
              const float Xc = fminf(X, 15.0f), Yc = fminf(Y, 15.0f);
              val = q == 0 ? Xc : (q == 1 ? Yc : Xc * Yc);
            } else {
              val = q == 0 ? X : (q == 1 ? Y : (own.x * own.y) * (a.x * a.y));
            }
          }
          const unsigned hit = __ballot_sync(0xffffffffu, val == target);
          if (hit) {
            const int src = __ffs(hit) - 1;
            f_off = __shfl_sync(0xffffffffu, off, src);
            f_x = __shfl_sync(0xffffffffu, X, src);
            f_y = __shfl_sync(0xffffffffu, Y, src);
            found = true;
          }
        }
        if (!found) continue;  // cannot happen (same arithmetic as phase 1); be safe
        const bool h1 = L.kind == kH1;
        const int opp_v = (int)((f_off - (h1 ? kp.tabQ1 : kp.tabP0)) / kRowBytes);
        const int own_v = (int)L.vteam;
        // weights of the two log-rates of the arg-max match (zero through a clipped rate)
        float wx = 0.0f, wy = 0.0f;
        if (which == 0) {
          const float wgt = gc * (1.0f - r) / Lam;  // dc/dLB * dLB/d eta
          if (q == 0) wx = (!CLIP || f_x < 15.0f) ? wgt : 0.0f;
          else wy = (!CLIP || f_y < 15.0f) ? wgt : 0.0f;
        } else {
          const float wgt = -gc * r / best[2];  // dc/dUB * dUB/d eta (UB = 1 / max lambda_h lambda_a)
          wx = (!CLIP || f_x < 15.0f) ? wgt : 0.0f;
          wy = (!CLIP || f_y < 15.0f) ? wgt : 0.0f;
        }
        const int ox = h1 ? eAh1 : eB0, oy = h1 ? eBh1 : eA0;  // own exponents of X, Y
        const int px = h1 ? eBa1 : eA0, py = h1 ? eAa1 : eB0;  // opponent's
#pragma unroll 1
        for (int side = 0; side < 2; side++) {
          const int sx = side ? px : ox, sy = side ? py : oy;
          float g[6];
#pragma unroll
          for (int e = 0; e < 6; e++) g[e] = (e == sx ? wx : 0.0f) + (e == sy ? wy : 0.0f);
          apply_vteam<KMAX>(kp, hy, acc, ln, lane == l, side ? opp_v : own_v, g, false, false, 0.0f, 0.0f, zero4);
        }
      }
    }
  }
  __syncthreads();  // tables are dead from here on; fix-ups of all chains are in memory

  // ---- confederation gradients -------------------------------------------------------------------------
  for (int k = warp; k < kp.Cf; k += W) {
    const float cf = ln.ld(o.conf + k);
    float s = -cf;  // N(0,1) prior
    acc.lp -= 0.5f * cf * cf;
    const int j0 = __ldg(kp.conf_vptr + k), j1 = __ldg(kp.conf_vptr + k + 1);
    for (int j = j0; j < j1; j++) s += ln.sc[(long long)__ldg(kp.conf_vlist + j) * kp.Cpad];
    if (ln.active) ln.gr[(long long)(o.conf + k) * ln.sd] = s;
  }

  // ---- cross-warp reduction of the hyper accumulators (table area is reused) ----------------------------
  const int NH = hyper_rows(kp.K);
  float* part = reinterpret_cast<float*>(smem);
  {
    float* p = part + (size_t)warp * NH * 32 + lane;
    p[0 * 32] = acc.lp; p[1 * 32] = acc.mu_d; p[2 * 32] = acc.ls_a; p[3 * 32] = acc.ls_d;
#pragma unroll
    for (int i = 0; i < 4; i++) { p[(4 + i) * 32] = acc.mu[i]; p[(8 + i) * 32] = acc.ls[i]; }
    p[12 * 32] = acc.rho;
    if (KMAX > 0) {
#pragma unroll
      for (int k = 0; k < KMAX; k++)
        if (k < kp.K) { p[(13 + k) * 32] = acc.ba[k]; p[(13 + kp.K + k) * 32] = acc.bd[k]; }
    }
  }
  __syncthreads();
  if (warp != 0) return;
  auto total = [&](int row) {
    float s = 0.0f;
    for (int w = 0; w < W; w++) s += part[((size_t)w * NH + row) * 32 + lane];
    return s;
  };
  float lp = total(0) + kp.const_term;
  auto gstore = [&](int d, float v) { if (ln.active) ln.gr[(long long)d * ln.sd] = v; };
  auto normal = [&](float x, float loc, float scale, int d, float extra) {
    const float z = (x - loc) / scale;
    lp += -0.5f * z * z - logf(scale) - kLogSqrt2Pi;
    gstore(d, -z / scale + extra);
  };
  auto halfnormal_exp = [&](float sig, float scale, int d, float extra) {  // HalfNormal on exp(x) + Jacobian x
    const float z = sig / scale;
    lp += -0.5f * z * z - logf(scale) - kLogSqrt2Pi + kLn2 + ln.ld(d);
    gstore(d, -z * z + 1.0f + extra);
  };
  const bool neu = kp.model == BPLX_NEUTRAL || kp.model == BPLX_NEUTRAL_WC;
  const float std_scale = neu ? 0.5f : 1.0f;  // neutral_dixon_coles.py:138-139
  normal(hy.mu_d, 0.0f, 1.0f, o.mean_defence, total(1));
  halfnormal_exp(hy.sig_a, std_scale, o.log_std_attack, total(2));
  halfnormal_exp(hy.sig_d, std_scale, o.log_std_defence, total(3));
  if (!neu) {
    normal(hy.mu[0], 0.1f, 0.2f, o.mean[0], total(4));
    if (ext) halfnormal_exp(hy.sig[0], 1.0f, o.log_std[0], total(8));
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      normal(hy.mu[i], (i & 1) ? -0.1f : 0.1f, 0.2f, o.mean[i], total(4 + i));
      halfnormal_exp(hy.sig[i], 1.0f, o.log_std[i], total(8 + i));
    }
  }
  if (KMAX > 0) {
#pragma unroll
    for (int k = 0; k < KMAX; k++) {
      if (k < kp.K) {
        normal(beta_a[k], 0.0f, 1.0f, o.beta_a + k, total(13 + k));
        normal(beta_d[k], 0.0f, 1.0f, o.beta_d + k, total(13 + kp.K + k));
      }
    }
  }
  // per-team constants of the standardised pair and the decentred sites
  lp -= (float)kp.T * (2.0f + (float)ndec) * kLogSqrt2Pi;
  lp -= (float)kp.Cf * kLogSqrt2Pi;
  if (has_rho) {  // u ~ Beta(2,4) + sigmoid Jacobian; rho = 2u - 1
    lp += 0.5f * (float)kp.T * logf(hy.inv_s2);
    const float lu = logf(u), l1u = logf(1.0f - u);
    lp += 2.0f * lu + 4.0f * l1u + 2.99573227355399099f;  // log 20
    gstore(o.u, 2.0f - 6.0f * u + total(12) * 2.0f * u * (1.0f - u));
  }
  {  // corr_coef_raw ~ Beta(2,2) + Jacobian; corr_coef = LB + r (UB - LB)
    lp += 2.0f * (logf(r) + logf(1.0f - r)) + 1.79175946922805500f;  // log 6
    gstore(o.raw, 2.0f * (1.0f - 2.0f * r) + gc * r * (1.0f - r) * (UB - LB));
  }
  if (ln.active) {
    kp.lp[chain] = lp;
    if (kp.corr_coef) kp.corr_coef[chain] = cc;
  }
}

// ---- host launcher --------------------------------------------------------------------------------------
template <bool CLIP, int KMAX>
static int launch_t(const KernelParams& kp, cudaStream_t stream, bool set_attr) {
  auto* fn = &logdensity_kernel<CLIP, KMAX>;
  if (set_attr) {
    BPLX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return BPLX_OK;
  }
  const int grid = (kp.C + kChains - 1) / kChains;
  fn<<<grid, kp.nwarps * 32, kp.smem_total, stream>>>(kp);
  BPLX_CUDA(cudaGetLastError());
  note_launch(1);
  return BPLX_OK;
}

static int dispatch(const KernelParams& kp, cudaStream_t stream, bool set_attr) {
  if (kp.clip) return kp.K > 0 ? launch_t<true, kMaxCov>(kp, stream, set_attr) : launch_t<true, 0>(kp, stream, set_attr);
  return kp.K > 0 ? launch_t<false, kMaxCov>(kp, stream, set_attr) : launch_t<false, 0>(kp, stream, set_attr);
}

int logdensity_set_attributes(const KernelParams& kp) { return dispatch(kp, nullptr, true); }
int launch_logdensity(const KernelParams& kp, cudaStream_t stream) { return dispatch(kp, stream, false); }

}  // namespace bplx
