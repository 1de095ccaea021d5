// logdensity_dynamic.cu -- K1d: log-density + gradient of the dynamic (random-walk) Dixon-Coles model.
//
// Replaces value_and_grad(potential_fn) of DynamicNeutralDixonColesMatchPredictor._model
// (bpl/dynamic_dixon_coles.py:63-247) with bpl/_util.py:17-93, for a batch of chains; lane = chain.
// Team strengths follow a random walk over gameweeks, attack[j] = attack[j-1] + z[j] * std_attack[j]
// (":192-218"; BPLX_FLAG_DYNAMIC_AS_WRITTEN reproduces the reference literally, where the walk never reaches
// the rates), and a match only pairs teams of its own gameweek.  D = 10 G + 2 + 2 K + 7 G T parameters per chain do not
// fit on chip, so theta is streamed: the CTA walks the gameweeks in step with ONE set of tables for the current
// gameweek (double-buffered), a warp owns teams (t -> warp t mod W), and every (gameweek, team) is completed in a single
// visit, so each gradient entry is written once and never re-read (plan_dynamic.inc has the stream layout):
//
//   forward pass   gameweeks ascending: the walk's prefix sums (kept per team, also written to the workspace), the
//                  gameweek's tables, and the HOME lists for the maxima of the corr_coef bounds (bpl/_util.py:17-31)
//   bounds         maxima over warps -> LB, UB, corr_coef
//   backward pass  gameweeks descending: tables again (from the stored prefix sums), then per team its rate lists
//                  (Poisson part), tau lists (bpl/_util.py:54-91), the running suffix sums (the walk's transpose),
//                  priors, chain rule and the final gradient of its seven sites; the ten per-gameweek hyper sums go
//                  through shared memory to one warp per gameweek.  The entry that attained each maximum is found
//                  while its gameweek is resident.
//   fix-up         d corr_coef / d eta of the two arg-max matches needs d/d corr_coef of ALL tau
//                  terms, known only now: it is linear, so it is applied as a correction to the few entries it reaches.
#include <math.h>

#include "k1_common.cuh"
#include "problem.h"

namespace bplx {

namespace {

// the pieces of a warp's stream, across stage boundaries (pieces never straddle a stage)
struct PieceReader {
  Ring* ring;
  uint32_t k, a, a0, aend, base;
  __device__ __forceinline__ void start(Ring* r, uint32_t base_) {
    ring = r;
    base = base_;
    k = 0;
    uint32_t bytes;
    a0 = ring->acquire(0, &bytes);
    a = a0;
    aend = a0 + bytes;
  }
  // next real header (fillers skipped); *hoff = its byte offset in the global stream; entries follow at `a`
  __device__ __forceinline__ Hdr next(uint32_t* hoff) {
    for (;;) {
      if (a + 16 > aend) {
        ring->release(k);
        k++;
        uint32_t bytes;
        a0 = ring->acquire(k, &bytes);
        a = a0;
        aend = a0 + bytes;
        continue;
      }
      const Hdr L = unpack_hdr(lds128u(a));
      if (L.flags & kStageEnd) {
        a = aend;
        continue;
      }
      *hoff = base + k * ring->S + (a - a0);
      a += 16;
      return L;
    }
  }
};

__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }  // written once, never re-read here

}  // namespace

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
  const uint32_t tab0 = smem_u32(smem) + lane * 8;
  float* const state = reinterpret_cast<float*>(smem + kp.dyn_state) + lane;  // [(t*2 + {att, def}) * 32]
  float* const part = reinterpret_cast<float*>(smem + kp.dyn_part) + lane;    // [((buf*W + warp)*10 + i) * 32]
  // per-gameweek hyper-parameters [(j*16 + i) * hs]: shared memory when it fits, else the workspace behind the prefix sums
  const int hs = kp.dyn_hyp_ws ? kp.Cpad : 32;
  float* const hyp = kp.dyn_hyp_ws ? kp.scratch + (size_t)G * T * 2 * kp.Cpad + chain
                                   : reinterpret_cast<float*>(smem + kp.dyn_hyp) + lane;
  float* const fin = reinterpret_cast<float*>(smem + kp.dyn_fin) + lane;      // [(t*8 + i) * 32] sites of the current gameweek
  unsigned long long* red_best = reinterpret_cast<unsigned long long*>(smem + kp.smem_red);  // [3][32]
  uint32_t* red_found = reinterpret_cast<uint32_t*>(smem + kp.smem_red + 768);               // [2][32]
  uint32_t* red_info = reinterpret_cast<uint32_t*>(smem + kp.smem_red + 1024);               // [2][32]
  float* red_gc = reinterpret_cast<float*>(smem + kp.smem_red + 1280);                        // [W][32]
  float* red_lp = red_gc + W * 32;                                                            // [W][32]
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
  const int nbuf = kp.dyn_nbuf;

  if (warp == 0) {
    const uint32_t zr = (uint32_t)T * kRowBytes;  // zero rows used by padding entries
    for (int b = 0; b < nbuf; b++) {
      const uint32_t tb = tab0 + b * kp.tab_bytes;
      if (kp.has1) {
        sts64(tb + kp.tabP1 + zr, 0.0f, 0.0f);
        sts64(tb + kp.tabQ1 + zr, 0.0f, 0.0f);
      }
      if (kp.has0) sts64(tb + kp.tabP0 + zr, 0.0f, 0.0f);
    }
    red_best[lane] = red_best[32 + lane] = red_best[64 + lane] = 0ull;
    red_found[lane] = red_found[32 + lane] = 0xffffffffu;
  }
  // the walk starts at the prior means: attack = X beta_a, defence = mean_defence + X beta_d
  for (int t = warp; t < T; t += W) {
    float a0 = 0.0f, d0 = mu_d;
    for (int k = 0; k < kp.K; k++) {
      const float x = __ldg(kp.Xs + (size_t)t * kp.K + k);
      a0 = fmaf(x, ln.ld(o.beta_a + k), a0);
      d0 = fmaf(x, ln.ld(o.beta_d + k), d0);
    }
    state[(t * 2) * 32] = a0;
    state[(t * 2 + 1) * 32] = d0;
  }
  // per-gameweek hyper-parameters, all computed up front (one round trip, gameweek j by warp j mod W):
  // [0..3] mu, [4..7] sig, [8] sig_attack, [9] sig_defence, [10..13] log sig, [14] log sig_attack, [15] log sig_defence
  struct GwHyp { float mu[4], sig[4], sig_a, sig_d; };
  auto read_hyp = [&](int j) {
    GwHyp h;
    const float* p = hyp + (size_t)j * 16 * hs;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      h.mu[i] = p[i * hs];
      h.sig[i] = p[(4 + i) * hs];
    }
    h.sig_a = p[8 * hs];
    h.sig_d = p[9 * hs];
    return h;
  };
  for (int j = warp; j < G; j += W) {
    float* h = hyp + (size_t)j * 16 * hs;
    float v[10];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      v[i] = ln.ld(o.mean[i] + j);
      v[4 + i] = ln.ld(o.log_std[i] + j);
    }
    v[8] = ln.ld(o.log_std_attack + j);
    v[9] = ln.ld(o.log_std_defence + j);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      h[i * hs] = v[i];
      h[(4 + i) * hs] = expf(v[4 + i]);
      h[(10 + i) * hs] = v[4 + i];
    }
    h[8 * hs] = expf(v[8]);
    h[9 * hs] = expf(v[9]);
    h[14 * hs] = v[8];
    h[15 * hs] = v[9];
  }
  __syncthreads();

  // table rows of (gameweek j, team t); with_lp: add the static sum of w * y * log(lambda) (linear in the exponents)
  auto build_row = [&](uint32_t tab, int t, int jt, float att, float def, const float (&dec)[4], const GwHyp& h,
                       bool with_lp) {
    float x[4];
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = fmaf(h.sig[i], dec[i], h.mu[i]);
    float ex[6];
    ex[eAh1] = att + x[0];
    ex[eBh1] = -def - x[2];
    ex[eBa1] = -def - x[3];
    ex[eAa1] = att + x[1];
    ex[eA0] = att;
    ex[eB0] = -def;
    const uint32_t row = (uint32_t)t * kRowBytes;
    // (the hardware exponential, 2 ulp: both passes build the rows with the same instructions, so the arg-max search
    //  still finds the very product the forward pass saw)
    if (kp.has1) {
      sts64(tab + kp.tabP1 + row, __expf(ex[eAh1]), __expf(ex[eBh1]));
      sts64(tab + kp.tabQ1 + row, __expf(ex[eBa1]), __expf(ex[eAa1]));
    }
    if (kp.has0) sts64(tab + kp.tabP0 + row, __expf(ex[eA0]), __expf(ex[eB0]));
    if (with_lp) {
#pragma unroll
      for (int e = 0; e < 6; e++) lp_acc = fmaf(__ldg(kp.yexp + (size_t)jt * 6 + e), ex[e], lp_acc);
    }
  };

  // ---- forward pass ---------------------------------------------------------------------------------------
  float best[3] = {0.0f, 0.0f, 0.0f};
  uint32_t besth[3] = {0u, 0u, 0u};
  PieceReader rdr;
  rdr.start(&ring, b1_0);
  for (int j = 0; j < G; j++) {
    const int hb = j & 1;
    const GwHyp h = read_hyp(j);
    const uint32_t tab = tab0 + (nbuf == 2 ? hb : 0) * kp.tab_bytes;
    for (int t = warp; t < T; t += W) {
      const int jt = j * T + t;
      float att = 0.0f, def = 0.0f;  // as written (dynamic_dixon_coles.py:192-218): the .at[].set results are discarded
      if (!kp.as_written) {
        att = fmaf(ln.ld(o.za + jt), h.sig_a, state[(t * 2) * 32]);
        def = fmaf(ln.ld(o.zd + jt), h.sig_d, state[(t * 2 + 1) * 32]);
        state[(t * 2) * 32] = att;
        state[(t * 2 + 1) * 32] = def;
        if (ln.active) {
          ln.sc[(size_t)(2 * jt) * kp.Cpad] = att;
          ln.sc[(size_t)(2 * jt + 1) * kp.Cpad] = def;
        }
      }
      if (__ldg(kp.team_flags + jt) & 1) {
        float dec[4];
#pragma unroll
        for (int i = 0; i < 4; i++) dec[i] = ln.ld(o.dec[i] + jt);
        build_row(tab, t, jt, att, def, dec, h, false);
      }
    }
    __syncthreads();  // the gameweek's tables are complete
    uint32_t hoff;
    const Hdr Mk = rdr.next(&hoff);
    for (uint32_t p = 0; p < Mk.n0; p++) {
      const Hdr L = rdr.next(&hoff);
      uint32_t a = rdr.a;
      const uint32_t e_end = a + L.n0 * ESZ;
      float2 own = lds64(tab + L.own_off);
      if (L.kind >= kH0) { const float s = own.x; own.x = own.y; own.y = s; }
      float m1 = 0.0f, m2 = 0.0f, m3 = 0.0f;
      walk16(a, e_end, [&](const uint32_t at) {
        const uint4 q = lds128u(at);  // two entries
        const float2 ea = lds64(tab + q.x), eb = lds64(tab + q.z);
        m1 = fmaxf(m1, fmaxf(ea.x, eb.x));
        m2 = fmaxf(m2, fmaxf(ea.y, eb.y));
        m3 = fmaxf(m3, fmaxf(ea.x * ea.y, eb.x * eb.y));
      });
      rdr.a = a;
      const float v0 = own.x * m1, v1 = own.y * m2, v2 = (own.x * own.y) * m3;
      if (v0 > best[0]) { best[0] = v0; besth[0] = hoff; }
      if (v1 > best[1]) { best[1] = v1; besth[1] = hoff; }
      if (v2 > best[2]) { best[2] = v2; besth[2] = hoff; }
    }
    if (nbuf == 1) __syncthreads();  // (one table buffer: nobody may still read it when the next gameweek is built)
  }
  const uint32_t b2_0 = __ldg(kp.warp_b2 + warp), b2_1 = __ldg(kp.warp_b2 + warp + 1);
  ring.begin(kp.stream2 + b2_0, b2_1 - b2_0);

  // ---- bounds (bpl/_util.py:17-31) ------------------------------------------------------------------
#pragma unroll
  for (int q = 0; q < 3; q++)
    if (best[q] > 0.0f)
      atomicMax(red_best + q * 32 + lane, ((unsigned long long)__float_as_uint(best[q]) << 32) | besth[q]);
  // the backward pass accumulates suffix sums in the per-team state
  for (int t = warp; t < T; t += W) state[(t * 2) * 32] = state[(t * 2 + 1) * 32] = 0.0f;
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
  // the two arg-max pieces and their gameweeks (0xffff: no dependence on the rates -- UB = 1)
  uint32_t s_hoff[2], s_gw[2];
  s_hoff[0] = qlam == 0 ? besth[0] : besth[1];
  s_hoff[1] = besth[2];
  s_gw[0] = Lam > 0.0f ? unpack_hdr(__ldg(reinterpret_cast<const uint4*>(kp.stream1 + s_hoff[0]))).vteam : 0xffffu;
  s_gw[1] = best[2] > 1.0f ? unpack_hdr(__ldg(reinterpret_cast<const uint4*>(kp.stream1 + s_hoff[1]))).vteam : 0xffffu;

  // The ten hyper-parameter sites of gameweek j: likelihood sums from the warps' partials + priors + Jacobians.  Component
  // c (0: std_attack, 1: std_defence, 2..5: mean_*, 6..9: std_*) is summed by warp c mod W, in a fixed order.
  auto reduce_hyper = [&](int j) {
    const float* pb = part + (size_t)((j & 1) * W) * 10 * 32;
    const float* hj = hyp + (size_t)j * 16 * hs;
    for (int c = warp; c < 10; c += W) {
      float s = 0.0f;
      for (int w = 0; w < W; w++) s += pb[(w * 10 + c) * 32];
      if (c < 2 || c >= 6) {  // HalfNormal(1) on exp(x) + Jacobian x
        const int i = c < 2 ? 8 + c : c - 2;        // index of sig; its log sits 6 entries further
        const float sig = hj[i * hs], lsig = hj[(i + 6) * hs];
        lp_acc += fmaf(-0.5f * sig, sig, lsig);
        const int site = c == 0 ? o.log_std_attack : (c == 1 ? o.log_std_defence : o.log_std[c - 6]);
        if (ln.active) st_stream(ln.g(site + j), fmaf(-sig, sig, 1.0f) + s);
      } else {  // mean_* ~ N(+-0.1, 0.2)
        const int i = c - 2;
        const float z = (hj[i * hs] - ((i & 1) ? -0.1f : 0.1f)) * 5.0f;
        lp_acc -= 0.5f * z * z;
        if (ln.active) st_stream(ln.g(o.mean[i] + j), fmaf(-z, 5.0f, s));
      }
    }
  };

  // ---- backward pass ----------------------------------------------------------------------------------
  float gc = 0.0f;
  rdr.start(&ring, b2_0);
  for (int j = G - 1; j >= 0; j--) {
    const int hb = j & 1;
    const GwHyp h = read_hyp(j);
    const uint32_t tab = tab0 + (nbuf == 2 ? hb : 0) * kp.tab_bytes;
    for (int t = warp; t < T; t += W) {
      // everything this (gameweek, team) needs from global memory is requested here, in one round trip; what the
      // completion step below needs waits in the warp's stash
      const int jt = j * T + t;
      const bool has = __ldg(kp.team_flags + jt) & 1;
      float att = 0.0f, def = 0.0f;
      if (has && !kp.as_written) {
        att = ld_cg(ln.sc + (size_t)(2 * jt) * kp.Cpad);
        def = ld_cg(ln.sc + (size_t)(2 * jt + 1) * kp.Cpad);
      }
      float dec[4];
#pragma unroll
      for (int i = 0; i < 4; i++) dec[i] = ln.ld(o.dec[i] + jt);
      const float za = ln.ld(o.za + jt), zd = ln.ld(o.zd + jt), ul = ln.ld(o.u + jt);
      float* f = fin + (size_t)(t * 8) * 32;
      f[0] = za;
      f[32] = zd;
      f[64] = ul;
#pragma unroll
      for (int i = 0; i < 4; i++) f[(3 + i) * 32] = dec[i];
      if (has) build_row(tab, t, jt, att, def, dec, h, true);
    }
    __syncthreads();  // tables complete; the hyper partials of gameweek j + 1 are all written
    if (j + 1 < G) reduce_hyper(j + 1);
    // arg-max search: the entry of the piece that attained the maximum, while its gameweek is resident; warp w looks
    // at entries w, w + W, ...; ties (rates that overflowed to inf) go to the lowest entry -- atomicMin, not timing
#pragma unroll 1
    for (int which = 0; which < 2; which++) {
      const bool need = s_gw[which] == (uint32_t)j;
      if (!__any_sync(kFull, need)) continue;
      const float target = which == 0 ? Lam : best[2];
      const int q = which == 0 ? qlam : 2;
      const unsigned char* base = kp.stream1 + (need ? s_hoff[which] : b1_0);
      const Hdr P = unpack_hdr(__ldg(reinterpret_cast<const uint4*>(base)));
      const uint32_t n = need ? P.n0 : 0u;
      float2 own = lds64(tab + (need ? P.own_off : 0u));
      if (P.kind >= kH0) { const float s = own.x; own.x = own.y; own.y = s; }
      const uint32_t nmax = __reduce_max_sync(kFull, n);
      uint32_t found = 0xffffffffu;
      for (uint32_t i = warp; i < nmax; i += W) {
        if (i < n && found == 0xffffffffu) {
          const uint32_t off = __ldg(reinterpret_cast<const uint32_t*>(base + 16 + (size_t)i * ESZ));
          const float2 ea = lds64(tab + off);
          const float val = q == 0 ? own.x * ea.x : (q == 1 ? own.y * ea.y : (own.x * own.y) * (ea.x * ea.y));
          if (val == target) found = (i << 24) | off;
        }
      }
      if (found != 0xffffffffu) {
        atomicMin(red_found + which * 32 + lane, found);
        red_info[which * 32 + lane] = P.team | (P.kind << 16);  // own team | kind: the same for every finder of this chain
      }
    }
    // the gameweek's pieces: every team of this warp is completed here
    float hp[10];
#pragma unroll
    for (int i = 0; i < 10; i++) hp[i] = 0.0f;
    float g[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    uint32_t hoff;
    const Hdr Mk = rdr.next(&hoff);
    for (uint32_t p = 0; p < Mk.n0; p++) {
      const Hdr L = rdr.next(&hoff);
      uint32_t a = rdr.a;
      if (L.flags & kTeamFirst) {
#pragma unroll
        for (int e = 0; e < 6; e++) g[e] = 0.0f;
      }
      float2 own = lds64(tab + L.own_off);
      if (L.kind >= kH0) { const float s = own.x; own.x = own.y; own.y = s; }
      const bool home = (L.kind & 1) == 0;
      if (!(L.flags & kPhase2)) {  // rates: sum w lambda over the list (every match is in two lists)
        const uint32_t e_end = a + L.n0 * ESZ;
        float2 acc0 = make_float2(0.0f, 0.0f), acc1 = acc0;
        walk16(a, e_end, [&](const uint32_t at) {
          const uint4 q = lds128u(at);
          const float2 ea = lds64(tab + q.x), eb = lds64(tab + q.z);
          acc0 = fma2(bc2(__uint_as_float(q.y)), ea, acc0);
          acc1 = fma2(bc2(__uint_as_float(q.w)), eb, acc1);
        });
        const float SX = own.x * (acc0.x + acc1.x), SY = own.y * (acc0.y + acc1.y);
        lp_acc -= 0.5f * (SX + SY);
        add_own(g, L.kind, -SX, -SY);
      } else {  // tau terms (bpl/_util.py:54-91)
        float lt, du, gx, gy;
        tau_piece<kTauPlain>(a, L, own, home, cc, tab, lt, du, gx, gy);
        if (home) {
          lp_acc = fmaf(lt, kLn2, lp_acc);
          gc += du;
        }
        add_own(g, L.kind, gx, gy);
      }
      rdr.a = a;
      if (L.flags & kTeamLast) {
        // ---- (gameweek j, team t) is complete: suffix sums, priors, chain rule, its seven gradient entries -----
        const int t = (int)L.team, jt = j * T + t;
        const float4 ys = __ldg(reinterpret_cast<const float4*>(kp.yteam + (size_t)jt * 8));
        const float2 ys2 = __ldg(reinterpret_cast<const float2*>(kp.yteam + (size_t)jt * 8 + 4));
        const float* f = fin + (size_t)(t * 8) * 32;
        const float za = f[0], zd = f[32], ul = f[64];
        float dec[4];
#pragma unroll
        for (int i = 0; i < 4; i++) dec[i] = f[(3 + i) * 32];
        const float ra = g[eAh1] + g[eAa1] + g[eA0] + ys.x;
        const float rd = -(g[eBh1] + g[eBa1] + g[eB0]) + ys.y;
        const float rx[4] = {g[eAh1] + ys.z, g[eAa1] + ys.w, -g[eBh1] + ys2.x, -g[eBa1] + ys2.y};
        float s_att = 0.0f, s_def = 0.0f;
        if (!kp.as_written) {  // d/d attack[j] reaches every later gameweek's rates: the walk's transpose
          s_att = state[(t * 2) * 32] + ra;
          s_def = state[(t * 2 + 1) * 32] + rd;
          state[(t * 2) * 32] = s_att;
          state[(t * 2 + 1) * 32] = s_def;
        }
        // u ~ Beta(2,4) + Jacobian; za ~ N(0,1); zd ~ N(rho za, sqrt(1 - rho^2))  (dynamic_dixon_coles.py:128-143)
        // (600+ of these per chain and call: the hardware's exp / log / reciprocal approximations, 2 ulp, are ample for
        //  a term of a sum checked at 1e-5 relative)
        const float u = fminf(fmaxf(__fdividef(1.0f, 1.0f + __expf(-ul)), FLT_MIN), 1.0f - FLT_EPSILON);
        const float rho = 2.0f * u - 1.0f, s2 = fmaf(-rho, rho, 1.0f), inv_s2 = __fdividef(1.0f, s2);
        const float e = zd - rho * za, es = e * inv_s2;
        lp_acc += -0.5f * (za * za + e * es) + kLn2 * (-0.5f * lg2_approx(s2) + 2.0f * lg2_approx(u) + 4.0f * lg2_approx(1.0f - u));
        const float a_rho = es * za - rho * es * es + rho * inv_s2;
        if (ln.active) {
          st_stream(ln.g(o.u + jt), 2.0f - 6.0f * u + a_rho * 2.0f * u * (1.0f - u));
          st_stream(ln.g(o.za + jt), fmaf(h.sig_a, s_att, -za + rho * es));
          st_stream(ln.g(o.zd + jt), fmaf(h.sig_d, s_def, -es));
        }
        hp[0] = fmaf(h.sig_a * za, s_att, hp[0]);
        hp[1] = fmaf(h.sig_d * zd, s_def, hp[1]);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          lp_acc -= 0.5f * dec[i] * dec[i];
          if (ln.active) st_stream(ln.g(o.dec[i] + jt), fmaf(h.sig[i], rx[i], -dec[i]));
          hp[2 + i] += rx[i];
          hp[6 + i] = fmaf(h.sig[i] * dec[i], rx[i], hp[6 + i]);
        }
      }
    }
    {
      float* pb = part + (size_t)((hb * W + warp) * 10) * 32;
#pragma unroll
      for (int i = 0; i < 10; i++) pb[i * 32] = hp[i];
    }
    if (nbuf == 1) __syncthreads();
  }
  red_gc[warp * 32 + lane] = gc;
  __syncthreads();  // partials of gameweek 0, the suffix sums of every team and every gradient entry are written
  reduce_hyper(0);
  gc = 0.0f;
  for (int w = 0; w < W; w++) gc += red_gc[w * 32 + lane];
  {  // the 1-1 matches: tau = 1 - c for all of them
    const float t11 = fmaxf(1.0f - cc, 0.0f);
    gc -= kp.w11 / t11;
    if (warp == 0 && kp.w11 != 0.0f) lp_acc = fmaf(kp.w11, logf(t11), lp_acc);
  }
  __syncthreads();  // (reduce_hyper(0) has written gameweek 0's hyper gradients: the fix-up below may add to them)

  // ---- arg-max fix-up, as a correction of the entries it reaches --------------------------
  Fixup fx;
  fx.h1 = 0u;
  fx.confs = 0u;
  int f_gw[2], f_team[2][2];
#pragma unroll
  for (int which = 0; which < 2; which++) {
    const uint32_t packed = red_found[which * 32 + lane];
    fx.teams[which] = fx.vts[which] = 0xffffffffu;
    fx.vx[which] = fx.vy[which] = 0.0f;
    f_gw[which] = -1;
    f_team[which][0] = f_team[which][1] = 0;
    if (packed != 0xffffffffu) {
      const uint32_t info = red_info[which * 32 + lane];
      const uint32_t f_off = packed & 0xffffffu;
      const bool h1 = ((info >> 16) & 3u) == kH1;
      if (which == 0) {
        const float wgt = gc * (1.0f - r) / Lam;  // dc/dLB * dLB/d eta
        fx.vx[0] = qlam == 0 ? wgt : 0.0f;
        fx.vy[0] = qlam == 1 ? wgt : 0.0f;
      } else {
        const float wgt = -gc * r / best[2];  // dc/dUB * dUB/d eta
        fx.vx[1] = fx.vy[1] = wgt;
      }
      fx.h1 |= (h1 ? 1u : 0u) << which;
      f_gw[which] = (int)s_gw[which];
      f_team[which][0] = (int)(info & 0xffffu);
      f_team[which][1] = (int)(((f_off & ~7u) - (h1 ? kp.tabQ1 : kp.tabP0)) / kRowBytes);
    }
  }
  float fix_md = 0.0f;  // what the fix-up adds to sum_t d/d defence[0, t] (mean_defence) -- and per team for the coefficients
  // the (which, side) items: team and what the fix-up adds to its raw (att, def, venue effect) gradients
  float f_dra[4], f_drd[4], f_drx[4][4];
#pragma unroll
  for (int it = 0; it < 4; it++) {
    f_dra[it] = f_drd[it] = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) f_drx[it][i] = 0.0f;
    if (f_gw[it >> 1] >= 0) fold_fixup(fx, it >> 1, it & 1, f_dra[it], f_drd[it], f_drx[it]);
  }
  for (int j = warp; j < G; j += W) {
    // all loads of the gameweek first (they are independent), then the updates: one round trip per gameweek instead of
    // one per entry; the scale factors come from the hyper-parameter table
    const float* hj = hyp + (size_t)j * 16 * hs;
    const float sa = hj[8 * hs], sd = hj[9 * hs];
    float za[4], zd[4], gza[4], gzd[4];
    bool on[4];
#pragma unroll
    for (int it = 0; it < 4; it++) {
      on[it] = f_gw[it >> 1] >= j && !kp.as_written && ln.active;
      const int jt = j * T + f_team[it >> 1][it & 1];
      za[it] = on[it] ? ln.ld(o.za + jt) : 0.0f;
      zd[it] = on[it] ? ln.ld(o.zd + jt) : 0.0f;
      gza[it] = on[it] ? *ln.g(o.za + jt) : 0.0f;
      gzd[it] = on[it] ? *ln.g(o.zd + jt) : 0.0f;
    }
    float d_lsa = 0.0f, d_lsd = 0.0f;
#pragma unroll
    for (int it = 0; it < 4; it++) {
      if (!on[it]) continue;
      // the two arg-max matches often share a team (the strongest attack is in both): one update per distinct team
      float dra = f_dra[it], drd = f_drd[it];
#pragma unroll
      for (int it2 = it + 1; it2 < 4; it2++) {
        if (on[it2] && f_team[it2 >> 1][it2 & 1] == f_team[it >> 1][it & 1]) {
          dra += f_dra[it2];
          drd += f_drd[it2];
          on[it2] = false;
        }
      }
      const int jt = j * T + f_team[it >> 1][it & 1];
      *ln.g(o.za + jt) = fmaf(sa, dra, gza[it]);
      *ln.g(o.zd + jt) = fmaf(sd, drd, gzd[it]);
      d_lsa = fmaf(sa * za[it], dra, d_lsa);
      d_lsd = fmaf(sd * zd[it], drd, d_lsd);
    }
    if (d_lsa != 0.0f || d_lsd != 0.0f) {
      *ln.g(o.log_std_attack + j) += d_lsa;
      *ln.g(o.log_std_defence + j) += d_lsd;
    }
#pragma unroll
    for (int it = 0; it < 4; it++) {
      if (f_gw[it >> 1] != j || !ln.active) continue;  // the venue effects of the match's own gameweek
      const int jt = j * T + f_team[it >> 1][it & 1];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float sig = hj[(4 + i) * hs];
        *ln.g(o.dec[i] + jt) += sig * f_drx[it][i];
        *ln.g(o.mean[i] + j) += f_drx[it][i];
        *ln.g(o.log_std[i] + j) += sig * ln.ld(o.dec[i] + jt) * f_drx[it][i];
      }
    }
  }
  // mean_defence and the covariate coefficients see the whole walk: sum_t d/d attack[0, t], d/d defence[0, t]
  float lp = 0.0f;
  const int wlast = W > 1 ? 1 : 0;
  if (warp == wlast) {
    float fdra[2][2], fdrd[2][2];
#pragma unroll
    for (int which = 0; which < 2; which++)
#pragma unroll
      for (int side = 0; side < 2; side++) {
        float dra = 0.0f, drd = 0.0f, drx[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (f_gw[which] >= 0 && !kp.as_written) fold_fixup(fx, which, side, dra, drd, drx);
        fdra[which][side] = dra;
        fdrd[which][side] = drd;
        fix_md += drd;
      }
    float s = 0.0f;
    if (!kp.as_written)
      for (int t = 0; t < T; t++) s += state[(t * 2 + 1) * 32];
    lp -= 0.5f * mu_d * mu_d;
    if (ln.active) *ln.g(o.mean_defence) = s + fix_md - mu_d;
    for (int k = 0; k < kp.K; k++) {
      float sa = 0.0f, sd = 0.0f;
      if (!kp.as_written) {
        for (int t = 0; t < T; t++) {
          const float x = __ldg(kp.Xs + (size_t)t * kp.K + k);
          sa = fmaf(x, state[(t * 2) * 32], sa);
          sd = fmaf(x, state[(t * 2 + 1) * 32], sd);
        }
#pragma unroll
        for (int which = 0; which < 2; which++)
#pragma unroll
          for (int side = 0; side < 2; side++) {
            const float x = __ldg(kp.Xs + (size_t)f_team[which][side] * kp.K + k);
            sa = fmaf(x, fdra[which][side], sa);
            sd = fmaf(x, fdrd[which][side], sd);
          }
      }
      const float ba = ln.ld(o.beta_a + k), bd = ln.ld(o.beta_d + k);
      lp -= 0.5f * (ba * ba + bd * bd);
      if (ln.active) {
        *ln.g(o.beta_a + k) = sa - ba;
        *ln.g(o.beta_d + k) = sd - bd;
      }
    }
    // corr_coef_raw ~ Uniform(0,1): Jacobian only; corr_coef = LB + r (UB - LB)
    lp += logf(r) + logf(1.0f - r);
    if (ln.active) {
      *ln.g(o.raw) = (1.0f - 2.0f * r) + gc * r * (1.0f - r) * (UB - LB);
      if (kp.corr_coef) kp.corr_coef[chain] = cc;
    }
  }
  red_lp[warp * 32 + lane] = lp + lp_acc;
  __syncthreads();
  if (warp == 0 && ln.active) {
    lp = kp.const_term;
    for (int w = 0; w < W; w++) lp += red_lp[w * 32 + lane];
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
