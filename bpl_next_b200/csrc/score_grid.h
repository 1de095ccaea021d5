// score_grid.h -- K3: posterior-predictive score grid (see score_grid.cu).
#pragma once
#include "common.cuh"

namespace bplx {

constexpr int kGridThreads = 256;   // fixtures per CTA (one thread per fixture)
constexpr int kGridStageMax = 16;   // posterior samples per TMA stage (fewer when a sample row is large)
constexpr int kGridStages = 2;      // stage buffers
constexpr int kGridStageBytes = 100 * 1024;  // budget of one stage buffer

struct GridParams {
  int model, S, T, Cf, F, g;
  int nsplit, samples_per_split;
  int ns_stage;    // samples per stage
  int ntab;        // per-team tables in a sample row: 2 (P1, Q1) or 3 (+ P0, models with neutral venues)
  int row_floats;  // floats per sample row of `table` (a multiple of 4: rows are 16-byte aligned for the bulk copies)
  int reuse_tables;  // the table in the workspace is already built for these samples
  float scale;
  // posterior samples, [S, T] row-major (home_advantage of DIXON_COLES: [S])
  const float *attack, *defence, *ha, *aa, *hd, *ad, *conf, *corr;
  const uint16_t *home, *away;
  const uint8_t *hconf, *aconf, *nv;
  float* table;    // [S][row_floats]  exponentials of the per-team log-rate halves (built by the pre-pass)
  float* partial;  // [nsplit][F][g*g]
  float* grid;     // [F][g*g]
  float* outcome;  // [F][3] or NULL
};

// fills nsplit / samples_per_split / ns_stage / ntab / row_floats of `gp` (S, T, Cf, F, g, model set) and returns the
// workspace bytes (table + partial sums); 0 with *err set when a sample row does not fit a stage buffer
size_t score_grid_plan(GridParams* gp, const char** err);
int launch_score_grid(const GridParams& gp, cudaStream_t stream);

}  // namespace bplx
