// nuts.h -- device-side view of include/bplx_nuts.h.
#pragma once
#include <stdint.h>

#include "../../include/bplx_nuts.h"

namespace bplx {

enum NutsStage : int { kNutsInitEval = 0, kNutsNewTransition = 1, kNutsInTree = 2, kNutsEvalPending = 3, kNutsDone = 4 };

struct NutsChain {  // scalar state of one chain
  int stage, t, window;
  float step_size, pe, energy_current;
  // trajectory
  int depth, going_right, turning, diverging, num_prop;
  float weight, sum_accept;
  // subtree under construction
  int sub_active, sub_num, sub_turning, sub_div;
  float sub_weight, sub_sum_accept, sub_pe;
  // dual averaging / Welford
  int da_t, wf_n;
  float da_mu, da_x, da_x_avg, da_g_avg;
  // counters
  int num_divergent;
  long long num_leapfrog_total;
  unsigned long long rng_offset;
};

struct NutsParams : bplx_nuts_params {
  __host__ __device__ NutsParams(const bplx_nuts_params& p) : bplx_nuts_params(p) {}
};

}  // namespace bplx
