// k1_common.cuh -- device helpers shared by the log-density kernels (logdensity.cu, logdensity_dynamic.cu).
#pragma once
#include <float.h>

#include "common.cuh"
#include "plan.h"

namespace bplx {

namespace {


constexpr float kLn2 = 0.693147180559945309f;
constexpr unsigned kFull = 0xffffffffu;

struct Hyp {  // constrained hyper-parameters of this lane's chain
  float mu_d, sig_a, sig_d, mu[4], sig[4];
};

struct Lane {
  const float* th;  // theta + chain * sc
  float* gr;        // grad  + chain * sc
  float* sc;        // scratch + chain
  int sd;
  bool active;
  __device__ __forceinline__ float ld(int d) const { return __ldg(th + (uint32_t)(d * sd)); }
  __device__ __forceinline__ float* g(int d) const { return gr + (uint32_t)(d * sd); }
};

__device__ __forceinline__ float sigmoid_clipped(float x) {
  // numpyro SigmoidTransform: clip(expit(x), finfo.tiny, 1 - finfo.eps)
  float s = 1.0f / (1.0f + expf(-x));
  return fminf(fmaxf(s, FLT_MIN), 1.0f - FLT_EPSILON);
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
__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }  // L2 only: the line is red.add'ed

// (two steps, so that a kernel can issue the loads early and do the exponentials when it needs the values)
__device__ __forceinline__ Hyp load_hyp_raw(const KernelParams& kp, const Lane& ln) {
  const ThetaOffsets& o = kp.off;
  Hyp hy;
  hy.mu_d = ln.ld(o.mean_defence);
  hy.sig_a = ln.ld(o.log_std_attack);
  hy.sig_d = ln.ld(o.log_std_defence);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    hy.mu[i] = o.mean[i] >= 0 ? ln.ld(o.mean[i]) : 0.0f;
    hy.sig[i] = o.log_std[i] >= 0 ? ln.ld(o.log_std[i]) : 0.0f;
  }
  return hy;
}
__device__ __forceinline__ Hyp finish_hyp(const KernelParams& kp, Hyp hy) {
  const ThetaOffsets& o = kp.off;
  hy.sig_a = expf(hy.sig_a);
  hy.sig_d = expf(hy.sig_d);
#pragma unroll
  for (int i = 0; i < 4; i++) hy.sig[i] = o.log_std[i] >= 0 ? expf(hy.sig[i]) : 0.0f;
  return hy;
}
__device__ __forceinline__ Hyp load_hyp(const KernelParams& kp, const Lane& ln) { return finish_hyp(kp, load_hyp_raw(kp, ln)); }

struct Hdr {
  uint32_t own_off, vteam, kind, flags, n0, n1, n2, team;
};
__device__ __forceinline__ Hdr unpack_hdr(const uint4 h) {
  Hdr H;
  H.own_off = h.x;
  H.vteam = h.y & 0xffffu;
  H.kind = (h.y >> 16) & 0xffu;
  H.flags = h.y >> 24;
  H.n0 = h.z & 0xffffu;
  H.n1 = h.z >> 16;
  H.n2 = h.w & 0xffffu;
  H.team = h.w >> 16;
  return H;
}

__device__ __forceinline__ void add_own(float (&g)[6], uint32_t kind, float gx, float gy) {
  if (kind == kH1) {
    g[eAh1] += gx; g[eBh1] += gy;
  } else if (kind == kA1) {
    g[eBa1] += gx; g[eAa1] += gy;
  } else {
    g[eB0] += gx; g[eA0] += gy;
  }
}

// exponent gradients of a team -> its raw slots (d/d att, d/d def, d/d venue effects).
// ADD = false: first touch (phase 1), plain stores.  ADD = true: red.add.
template <bool ADD>
__device__ __forceinline__ void put_raw(const KernelParams& kp, const Lane& ln, int t, const float (&g)[6], float& hacc) {
  const ThetaOffsets& o = kp.off;
  const float ra = g[eAh1] + g[eAa1] + g[eA0];
  const float rd = -(g[eBh1] + g[eBa1] + g[eB0]);
  const float rx[4] = {g[eAh1], g[eAa1], -g[eBh1], -g[eBa1]};
  if (kp.ndec == 0) hacc += rx[0];  // DIXON_COLES: scalar home advantage
  if (!ln.active) return;
  if (ADD) {
    red_add(ln.g(o.za + t), ra);
    red_add(ln.g(o.zd + t), rd);
  } else {
    *ln.g(o.za + t) = ra;
    *ln.g(o.zd + t) = rd;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (i < kp.ndec) {
      if (ADD) red_add(ln.g(o.dec[i] + t), rx[i]);
      else *ln.g(o.dec[i] + t) = rx[i];
    }
  }
}

// the same sums of one side's tau lists, stored to that side's shared-memory slots (slot i of a chain at q[i * 32])
__device__ __forceinline__ void put_side(const KernelParams& kp, float* q, const float (&g)[6], float& hacc) {
  q[0] = g[eAh1] + g[eAa1] + g[eA0];
  q[32] = -(g[eBh1] + g[eBa1] + g[eB0]);
  const float rx[4] = {g[eAh1], g[eAa1], -g[eBh1], -g[eBa1]};
  if (kp.ndec == 0) hacc += rx[0];
#pragma unroll
  for (int i = 0; i < 4; i++)
    if (i < kp.ndec) q[(2 + i) * 32] = rx[i];
}

// Per-warp TMA ring over a contiguous global byte stream (the warp's list pieces of one phase).
// One elected lane issues cp.async.bulk copies of one stage into the warp's private ring and every
// lane waits on the stage's mbarrier before reading it; the same warp produces and consumes, so a
// __syncwarp() is all that is needed before a slot is refilled.  Pieces never straddle a stage.
struct Ring {
  uint32_t ring, bar;        // shared addresses: kStages stages, kStages mbarriers
  uint32_t S;                // stage bytes
  const unsigned char* src;  // current stream
  uint32_t total;            // bytes in the current stream
  uint32_t gs0, gs_next;     // ring-stage counter at the start of / after the current stream
  int lane;

  __device__ __forceinline__ void init(uint32_t ring_, uint32_t bar_, uint32_t S_, int lane_) {
    ring = ring_; bar = bar_; S = S_; lane = lane_;
    total = gs0 = gs_next = 0;
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
    const uint32_t b0 = k * S;
    if (lane == 0 && b0 < total) {
      const uint32_t bytes = min(S, total - b0);
      const uint32_t slot = (gs0 + k) % kStages;
      mbar_arrive_expect_tx(bar + 8 * slot, bytes);
      tma_load_1d(ring + slot * S, src + b0, bytes, bar + 8 * slot);
    }
  }
  __device__ __forceinline__ void begin(const void* src_, uint32_t total_bytes) {
    __syncwarp();  // every lane is done with the previous stream's stages
    src = static_cast<const unsigned char*>(src_);
    total = total_bytes;
    gs0 = gs_next;
    gs_next = gs0 + (total_bytes + S - 1) / S;
#pragma unroll
    for (int s = 0; s < kStages; s++) issue(s);
  }
  __device__ __forceinline__ uint32_t num_stages() const { return (total + S - 1) / S; }
  // waits for stage k; returns its shared address, *bytes = its size
  __device__ __forceinline__ uint32_t acquire(uint32_t k, uint32_t* bytes) {
    const uint32_t idx = gs0 + k, slot = idx % kStages;
    mbar_wait(bar + 8 * slot, (idx / kStages) & 1);
    *bytes = min(S, total - k * S);
    return ring + slot * S;
  }
  __device__ __forceinline__ void release(uint32_t k) {
    __syncwarp();
    issue(k + kStages);
  }
};

// the two arg-max matches of a chain, as every warp needs them in the team pass
struct Fixup {
  uint32_t teams[2];  // per which: own team | opp team << 16 (0xffff = none)
  uint32_t confs;     // 4 x 8 bits: (which, side) -> confederation (WC)
  uint32_t vts[2];    // per which: own vteam | opp vteam << 16
  float vx[2], vy[2]; // gradient reaching the X / Y log-rate of the arg-max match
  uint32_t h1;        // bit which: the list kind is H1 (else H0)
};

// d/d (att, def, venue effects) of team t gets the fix-up of (which, side) when it is that match's team
__device__ __forceinline__ void fold_fixup(const Fixup& fx, int which, int side, float& ra, float& rd, float (&rx)[4]) {
  const bool h1 = (fx.h1 >> which) & 1u;
  const float vx = fx.vx[which], vy = fx.vy[which];
  if (side == 0) {
    if (h1) { ra += vx; rx[0] += vx; rd -= vy; rx[2] -= vy; }  // X -> A_h1, Y -> B_h1
    else { rd -= vx; ra += vy; }                                 // X -> B_0,  Y -> A_0
  } else {
    if (h1) { rd -= vx; rx[3] -= vx; ra += vy; rx[1] += vy; }  // X -> B_a1, Y -> A_a1
    else { ra += vx; rd -= vy; }                                 // X -> A_0,  Y -> B_0
  }
}
__device__ __forceinline__ float fixup_conf(const Fixup& fx, int which, int side) {  // d/d (A - B)
  const bool h1 = (fx.h1 >> which) & 1u;
  const float d = fx.vx[which] - fx.vy[which];
  return ((side == 0) == h1) ? d : -d;
}

}  // namespace

}  // namespace bplx
