// common.cuh -- error plumbing and sm_100a PTX helpers shared by the bplx kernels.
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/bplx.h"

namespace bplx {

// ---- thread-local error message -----------------------------------------------------------
void set_error(const char* fmt, ...);
void note_launch(int n = 1);  // kernels enqueued by this library (bplx_launch_count)

#define BPLX_CUDA(call)                                                                        \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      ::bplx::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return BPLX_E_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

#define BPLX_REQUIRE(cond, code, ...) \
  do {                                \
    if (!(cond)) {                    \
      ::bplx::set_error(__VA_ARGS__); \
      return (code);                  \
    }                                 \
  } while (0)

// Testing / tuning switches from the environment, read ONCE (first use) and again only by bplx_reload_env() -- never per
// launch:  BPLX_NO_PDL=1 (no programmatic dependent launch on the K1 / NUTS-step launches), BPLX_SPLIT=1|2|4|8 (CTAs
// per chain group), BPLX_HOST_CHUNKS=1|2|4|8|16 (host entry point pipeline), BPLX_NUTS_GENERIC=1 (stage-by-stage step kernel),
// BPLX_NO_TAIL_SPLIT=1 (no second, cluster-split launch for the last partial wave of chain groups),
// BPLX_NO_TRANSPOSE=1 (large [chains, D] batches are computed as they come, not through the native [D, chains] layout).
struct EnvSwitches {
  bool no_pdl = false, nuts_generic = false, no_tail_split = false, no_transpose = false;
  int split = 0, host_chunks = 0;
  size_t transpose_min_elems = (size_t)1 << 17;  // BPLX_TRANSPOSE_MIN_ELEMS: smallest chains x D device batch that is transposed
};
const EnvSwitches& env_switches();
void reload_env_switches();
inline bool pdl_enabled() { return !env_switches().no_pdl; }

// ---- device helpers -----------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// packed FP32 pairs (sm_100: FFMA2 / FMUL2 / FADD2 -- two IEEE fp32 operations per issue slot; a pair whose halves are
// the same register is taken as a broadcast scalar operand, a negated one folds into the instruction)
__device__ __forceinline__ unsigned long long pk2(float2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 upk2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c)));
  return upk2(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)));
  return upk2(d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)));
  return upk2(d);
}
__device__ __forceinline__ float2 bc2(float x) { return make_float2(x, x); }

// programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// while its predecessor in the stream is still running; everything it does before pdl_wait() must touch nothing the
// predecessor writes or reads-then-overwrites.  pdl_wait() returns when the predecessor has completed and its writes
// are visible; pdl_launch_dependents() lets the next kernel in the stream begin its own preamble.  Both are no-ops
// for ordinary launches.  Pointers to data the predecessor produces are passed through after_wait() so that no load
// from them (not even a read-only-path load the compiler considers movable) can be scheduled above the wait.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename T>
__device__ __forceinline__ T* after_wait(T* p) {
  asm volatile("" : "+l"(p));
  return p;
}

// distributed shared memory (thread-block clusters): 32-bit shared::cluster addresses ----------------------
// (generic-pointer atomics on a mapped address compile to a LOCAL shared-memory CAS loop -- address the remote CTA
//  explicitly.  A launch without a cluster attribute is an implicit cluster of one: rank 0 is the CTA itself.)
__device__ __forceinline__ uint32_t cluster_map(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void dsm_st_f32(uint32_t a, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void dsm_st_u32(uint32_t a, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ float dsm_ld_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t dsm_ld_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long dsm_ld_u64(uint32_t a) {
  unsigned long long v;
  asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void dsm_st_u64(uint32_t a, unsigned long long v) {
  asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}
// 32-bit max is a native shared-memory atomic for local and remote requesters alike (the 64-bit one is a local CAS loop
// on one side and a remote atomic on the other, which do not exclude each other: measured lost updates)
__device__ __forceinline__ void dsm_atom_min_u32(uint32_t a, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared::cluster.min.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void dsm_atom_max_u32(uint32_t a, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared::cluster.max.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
}

// mbarrier (shared::cta) -------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(gmem_src), "r"(bytes), "r"(bar)
               : "memory");
}
#endif  // __CUDACC__

}  // namespace bplx
