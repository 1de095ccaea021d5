// peer_sum.cu -- the one exchange of the path, over peer memory: the sum over ranks of the partial score grids.
//
// predict_* averages the score grid over the posterior samples (bpl/base.py:94-110, `.mean(axis=0)`); with the samples
// sharded over the GPUs every rank holds a partial grid already scaled by 1 / S_total, and the mean is their sum.  The
// ranks keep their partial grids in SYMMETRIC allocations (same size, every rank's buffer mapped into every other rank's
// address space over NVLink / NVSwitch -- torch.distributed._symmetric_memory hands out the peer pointers).  Per fixture
// range and rank: one warp publishes "my part of epoch e is written" into every peer's flag array and waits for the
// peers' flags in its own; the sum kernel behind it reads the same range of every rank's buffer -- rank 0 first, so every
// rank computes the same bits -- and writes the sum to its local result.  No NCCL call, no staging copy, no ring: the
// exchange of a 1.2 MB range is latency-bound, and this is two small launches instead of a collective.
#include <stdint.h>

#include "common.cuh"
#include "../../include/bplx.h"

namespace bplx {

namespace {

constexpr int kMaxPeers = 16;

struct PeerArgs {
  const float* data[kMaxPeers];  // every rank's buffer (data part), this rank's own included
  uint32_t* flags[kMaxPeers];    // every rank's flag array: flags[q][p] = last epoch rank p has published to rank q
  int n, rank;
  uint32_t epoch;
  size_t off[2], cnt[2];  // two element ranges of the buffers (the grid rows and the outcome rows of a fixture range)
  float* out[2];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// one warp: publish this rank's epoch to every peer, then wait for theirs.  Kept apart from the sum so that no CTA of
// the (wide) sum kernel ever spins on a slower rank next to this rank's own grid kernel of the following range.
__global__ void __launch_bounds__(32) peer_flag_kernel(const PeerArgs a) {
  const int lane = threadIdx.x;
  // epoch 0: the call counter kept behind the per-rank flags of this rank's own buffer (slot kMaxPeers) -- every rank makes
  // the same sequence of calls, so the counters agree, and a launch that needs no host argument can be replayed in a graph
  uint32_t epoch = a.epoch;
  if (epoch == 0u) {
    uint32_t* counter = a.flags[a.rank] + kMaxPeers;
    if (lane == 0) {
      epoch = *counter + 1u;
      if (epoch == 0u) epoch = 1u;
      *counter = epoch;
    }
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
  }
  __threadfence_system();  // what this rank wrote before (its partial grid) is visible to whoever sees the flag
  if (lane < a.n) st_release_sys(a.flags[lane] + a.rank, epoch);
  if (lane < a.n)  // (epochs only grow; the difference is taken so that a wrapped counter still works)
    while ((int32_t)(ld_acquire_sys(a.flags[a.rank] + lane) - epoch) < 0) {
    }
  __syncwarp();
  __threadfence_system();
}

__global__ void __launch_bounds__(512) peer_sum_kernel(const PeerArgs a) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const size_t off = a.off[s], cnt = a.cnt[s];
    float* out = a.out[s];
    if (cnt == 0) continue;
    const bool vec = (off % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
    const size_t n4 = vec ? cnt / 4 : 0;
    for (size_t i = tid; i < n4; i += nthr) {  // rank order: every rank adds the same numbers in the same order
      float4 v[kMaxPeers];
#pragma unroll
      for (int q = 0; q < kMaxPeers; q++)
        if (q < a.n) v[q] = __ldcv(reinterpret_cast<const float4*>(a.data[q] + off) + i);
      float4 acc = v[0];
#pragma unroll
      for (int q = 1; q < kMaxPeers; q++)
        if (q < a.n) {
          acc.x += v[q].x;
          acc.y += v[q].y;
          acc.z += v[q].z;
          acc.w += v[q].w;
        }
      reinterpret_cast<float4*>(out)[i] = acc;
    }
    for (size_t i = n4 * 4 + tid; i < cnt; i += nthr) {
      float acc = __ldcv(a.data[0] + off + i);
      for (int q = 1; q < a.n; q++) acc += __ldcv(a.data[q] + off + i);
      out[i] = acc;
    }
  }
}

}  // namespace
}  // namespace bplx

using namespace bplx;

extern "C" int bplx_peer_sum(void* const* bufs, int nranks, int rank, size_t flag_bytes, size_t off0, size_t cnt0,
                             float* out0, size_t off1, size_t cnt1, float* out1, unsigned epoch, void* stream) {
  BPLX_REQUIRE(bufs && nranks >= 1 && nranks <= kMaxPeers && rank >= 0 && rank < nranks, BPLX_E_INVALID,
               "peer_sum: bad ranks (%d of %d, at most %d)", rank, nranks, kMaxPeers);
  BPLX_REQUIRE(flag_bytes >= (size_t)(kMaxPeers + 1) * sizeof(uint32_t) && flag_bytes % 16 == 0, BPLX_E_INVALID,
               "peer_sum: the flag area must hold %d 32-bit words and keep the data 16-byte aligned", kMaxPeers + 1);
  BPLX_REQUIRE((cnt0 == 0 || out0) && (cnt1 == 0 || out1), BPLX_E_INVALID, "peer_sum: output is NULL");
  PeerArgs a{};
  for (int q = 0; q < nranks; q++) {
    BPLX_REQUIRE(bufs[q] != nullptr, BPLX_E_INVALID, "peer_sum: buffer of rank %d is NULL", q);
    a.flags[q] = static_cast<uint32_t*>(bufs[q]);
    a.data[q] = reinterpret_cast<const float*>(static_cast<const char*>(bufs[q]) + flag_bytes);
  }
  a.n = nranks;
  a.rank = rank;
  a.epoch = epoch;
  a.off[0] = off0, a.cnt[0] = cnt0, a.out[0] = out0;
  a.off[1] = off1, a.cnt[1] = cnt1, a.out[1] = out1;
  const size_t work = (cnt0 + cnt1 + 3) / 4;
  int blocks = (int)((work + 511) / 512);
  blocks = blocks < 1 ? 1 : (blocks > 64 ? 64 : blocks);
  peer_flag_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
  BPLX_CUDA(cudaGetLastError());
  peer_sum_kernel<<<blocks, 512, 0, static_cast<cudaStream_t>(stream)>>>(a);
  BPLX_CUDA(cudaGetLastError());
  note_launch(2);
  return BPLX_OK;
}
