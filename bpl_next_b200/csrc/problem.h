// problem.h -- the opaque bplx_problem handle.
#pragma once
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "plan.h"

struct bplx_problem {
  int device = 0;
  bplx::KernelParams kp{};  // static part filled at create; call arguments filled per launch
  std::string layout;
  // likelihood-only view of the same plan (bplx_loglik_fwdbwd): offsets / sites of the packed per-team tables
  bplx::ThetaOffsets lik_off{};
  bplx::HyperDesc lik_hyper[12]{};
  int lik_nhyper = 0, lik_D = 0;
  float lik_const = 0.0f;
  std::string lik_layout;
  // streams per split (1, 2, 4, 8 CTAs per chain group) and how many clusters of each size the device can hold at once
  const unsigned char* s1[bplx::kNumSplits] = {};
  const unsigned char* s2[bplx::kNumSplits] = {};
  const uint32_t* wb1[bplx::kNumSplits] = {};
  const uint32_t* wb2[bplx::kNumSplits] = {};
  bplx::WarpBounds wb[bplx::kNumSplits] = {};  // host copies of wb1 / wb2 (static models): passed by value at launch
  int max_clusters[bplx::kNumSplits] = {0, 0, 0, 0};
  std::vector<void*> dev_allocs;
  // host-variant staging (lazily grown, guarded by mu)
  std::mutex mu;
  int host_cap = 0;
  float *h_theta = nullptr, *h_out = nullptr;  // pinned
  float *d_theta = nullptr, *d_out = nullptr;
  float *d_tin = nullptr, *d_tout = nullptr;    // one chunk of theta / grad in the kernel's native chain-minor layout
  int tr_cap = 0;                               // ... their capacity in chains
  void* d_ws = nullptr;
  size_t d_ws_bytes = 0;
  cudaStream_t host_stream = nullptr;            // compute
  cudaStream_t host_in = nullptr, host_out = nullptr;  // H2D / D2H copies of the host variant (pipelined by chunk)
  cudaEvent_t host_ev[48] = {};                     // [2 * chunk]: chunk uploaded, chunk computed
  // plan statistics (for DESIGN/bench reporting)
  int sm_count = 0;
  long long stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<long long> warp_stats;  // [warps][12], see bplx_problem_warp_stats
};

namespace bplx {
// pdl: launch with programmatic stream serialization (device-pointer entry point: the predecessor in the stream is
// usually the kernel that produced theta; not after the host variant's copies)
int launch_logdensity(const KernelParams& kp, const WarpBounds& wb, cudaStream_t stream, bool pdl);
int logdensity_set_attributes(const KernelParams& kp);
int logdensity_max_clusters(const KernelParams& kp, int split);  // co-resident clusters of `split` CTAs, 0 if unsupported
int launch_logdensity_dynamic(const KernelParams& kp, cudaStream_t stream);
int logdensity_dynamic_set_attributes();
}  // namespace bplx
