/*
 * bplx_nuts.h -- C ABI of the batched NUTS transition kernel (the caller of bplx_logdensity_fwdbwd).
 *
 * Replaces, for this path, numpyro's `NUTS` + `MCMC(chain_method="vectorized")` as constructed in every reference
 * `fit` (bpl/dixon_coles.py:100-116, bpl/extended_dixon_coles.py:293-316, bpl/neutral_dixon_coles.py:330-346,
 * bpl/neutral_dixon_coles_WC.py:281-300, bpl/dynamic_dixon_coles.py:279-296): numpyro defaults (target accept 0.8,
 * max tree depth 10, diagonal mass, step size 1.0 adapted by dual averaging, Stan windows, divergence at dH > 1000).
 *
 * Protocol (all buffers caller-owned DEVICE memory, float32, chain-minor [D][ld] unless stated or state_layout == 1):
 *     bplx_nuts_init(&p)                          chains start at stage "evaluate the initial position"
 *     copy the initial positions into theta_eval
 *     repeat { potential: lp, grad <- log-density(theta_eval);  bplx_nuts_step(&p) }  until *active_count == 0
 * `active_count` is incremented by every chain that is not finished after a step: zero it before the step whose
 * count you want to read.  Chains are not in lock step: each starts its next transition as soon as its tree ends.
 */
#ifndef BPLX_NUTS_H_
#define BPLX_NUTS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int32_t start, end; } bplx_window; /* adaptation window [start, end] in transitions */

typedef struct bplx_nuts_params {
  int32_t C, D, ld;              /* chains, parameters, leading dimension (>= C) of every [D][ld] array */
  int32_t num_warmup, num_samples, thin, num_keep; /* num_keep = ceil(num_samples / thin) draws are stored */
  int32_t max_tree_depth;        /* numpyro default 10 */
  int32_t num_windows;
  const bplx_window* windows;    /* device, [num_windows] (Stan schedule, see bpl_next_b200/nuts.py) */
  float target_accept;           /* 0.8 */
  float init_step_size;          /* 1.0 */
  float max_delta_energy;        /* 1000 */
  uint64_t seed;
  int64_t chain_offset;          /* global index of chain 0 of this rank (decorrelates ranks) */
  /* per launch: input from the log-density call, output for the next one */
  float* theta_eval;             /* [D][ld] in/out */
  const float* lp;               /* [C] log density at theta_eval */
  float* grad;                   /* [D][ld] d lp / d theta at theta_eval */
  /* state */
  void* chain;                   /* [C] x bplx_nuts_chain_bytes() */
  float *p_half, *inv_mass, *zL, *rL, *gL, *zR, *rR, *gR, *zP, *gP, *r_sum, *zQ, *gQ, *r_sum_sub; /* [D][ld] each */
  float *r_ckpts, *r_sum_ckpts;  /* [max_tree_depth][D][ld] */
  float *wf_mean, *wf_m2;        /* [D][ld] Welford accumulators */
  /* output */
  float* samples;                /* [num_keep][D][ld] unconstrained draws */
  float* sample_lp;              /* [num_keep][ld] */
  float* sample_accept;          /* [num_keep][ld] mean acceptance probability of the draw's tree */
  int32_t* active_count;         /* device scalar */
  /* streaming diagnostics (optional: diag_lags == 0 turns them off).  Per chain and parameter, over the post-warm-up
   * draws x_0 .. x_{N-1} shifted by the chain's first draw (v_k = x_k - x_0), updated when a chain collects a draw:
   *   dg_sums[0..5] = sum v, sum v^2 over the whole chain, over the first N/2 draws, over the last N/2 draws
   *   dg_lag[l-1]   = sum_k v_k v_{k-l},  l = 1 .. diag_lags
   *   dg_ring       = the last diag_lags values (slot k mod diag_lags),  dg_head = the first diag_lags values
   * Split R-hat and the bulk ESS (autocovariances up to lag diag_lags) follow from these and sums over chains
   * (bpl_next_b200/diagnostics.py: streaming_summary), so no draw has to be stored to diagnose a run. */
  int32_t diag_lags;
  float* dg_ref;                 /* [D][ld]              x_0 */
  float* dg_sums;                /* [6][D][ld]           zero-initialised by the caller */
  float *dg_lag, *dg_ring, *dg_head; /* [diag_lags][D][ld] each, zero-initialised by the caller */
  /* Layout of every per-(chain, parameter) array above -- theta_eval and grad included:
   *   0  chain-minor [..][D][ld] (lane = chain; what the register-resident kernels for D <= 256 use)
   *   1  chain-major [..][C][ld_state], ld_state >= D (a warp per chain: for large D, where chains that are not in lock
   *      step make the chain-minor step touch every line for a few chains at a time -- 11.2 -> ~3 GB of DRAM traffic per
   *      step on configs[2]).  The log-density call then takes BPLX_CHAIN_MAJOR buffers with ld = ld_state.
   * The per-chain scalars (lp, sample_lp, sample_accept) keep the pitch ld in both. */
  int32_t state_layout, ld_state;
} bplx_nuts_params;

size_t bplx_nuts_chain_bytes(void);
int bplx_nuts_init(const bplx_nuts_params* p, void* stream);
int bplx_nuts_step(const bplx_nuts_params* p, void* stream);
/* per-chain summary after (or during) a run: out[c*8 + {0: transitions done, 1: step size, 2: divergences after warmup,
 * 3: leapfrogs total, 4: tree depth of the last transition, 5..7: reserved}] (host array of 8*C floats) */
int bplx_nuts_summary(const bplx_nuts_params* p, float* out_host);

#ifdef __cplusplus
}
#endif
#endif /* BPLX_NUTS_H_ */
