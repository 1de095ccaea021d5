// score_grid.h -- K3: posterior-predictive score grid (see score_grid.cu).
#pragma once
#include "common.cuh"

namespace bplx {

constexpr int kGridThreads = 256;  // fixtures per CTA (one thread per fixture)
constexpr int kGridStage = 8;      // posterior samples staged per cp.async stage

struct GridParams {
  int model, S, T, Cf, F, g;
  int nsplit, samples_per_split;
  float scale;
  // posterior samples, [S, T] row-major (home_advantage of DIXON_COLES: [S])
  const float *attack, *defence, *ha, *aa, *hd, *ad, *conf, *corr;
  const uint16_t *home, *away;
  const uint8_t *hconf, *aconf, *nv;
  float* partial;  // [nsplit][F][g*g]
  float* grid;     // [F][g*g]
  float* outcome;  // [F][3] or NULL
};

size_t score_grid_workspace(int S, int F, int g, int* nsplit, int* samples_per_split);
int launch_score_grid(const GridParams& gp, cudaStream_t stream);

}  // namespace bplx
