/*
 * bplx.h -- C ABI of the B200-native Dixon-Coles hot path for bpl-next.
 *
 * The reference (anguswilliams91/bpl-next) has no FFI: its hot path is the numpyro model callable
 * handed to NUTS (bpl/dixon_coles.py:100, bpl/extended_dixon_coles.py:293,
 * bpl/neutral_dixon_coles.py:330, bpl/neutral_dixon_coles_WC.py:281,
 * bpl/dynamic_dixon_coles.py:279) and the eager predict_* methods (bpl/base.py:74-148).  These
 * entry points are what a jax.ffi / ctypes binding for that path binds (INTEGRATION.md shows both).
 *
 * Conventions
 *   - plain C: pointers + sizes, int return (0 = ok, <0 = bplx_status), no exceptions, no exit();
 *     bplx_last_error() gives a thread-local message for the last failing call on this thread.
 *   - device pointers are caller-owned and must stay valid until the enqueued work completes;
 *     *_fwdbwd and bplx_score_grid are asynchronous: they enqueue on `stream` (a cudaStream_t
 *     passed as void*; NULL = legacy default stream), never synchronise and never allocate.
 *   - the *_host variants take HOST buffers, do the H2D/D2H copies themselves through pinned
 *     staging owned by the handle and return after the results are in the host buffers.
 *   - a bplx_problem is immutable after create and bound to the CUDA device that was current at
 *     create time; calls may come from any host thread (one stream per concurrent caller).
 *   - all floating point is IEEE binary32 (the reference runs JAX with x64 disabled); indices are
 *     the reference's DTYPES (bpl/base.py:16-22): uint16 teams, uint8 goals / conferences / venue.
 */
#ifndef BPLX_H_
#define BPLX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPLX_VERSION 1

typedef enum {
  BPLX_OK = 0,
  BPLX_E_INVALID = -1,     /* bad argument (null pointer, index out of range, ...) */
  BPLX_E_UNSUPPORTED = -2, /* shape outside what the kernels were built for */
  BPLX_E_CUDA = -3,        /* CUDA runtime error (message has the cudaError string) */
  BPLX_E_NOMEM = -4,
  BPLX_E_WORKSPACE = -5    /* workspace too small */
} bplx_status;

/* which bpl-next class the problem mirrors */
typedef enum {
  BPLX_DIXON_COLES = 0, /* bpl/dixon_coles.py:39-84          D = 5 + 2T            */
  BPLX_EXTENDED = 1,    /* bpl/extended_dixon_coles.py:78-248 D = 7 + 2K + 3T       */
  BPLX_NEUTRAL = 2,     /* bpl/neutral_dixon_coles.py:102-283 D = 13 + 2K + 6T      */
  BPLX_NEUTRAL_WC = 3,  /* bpl/neutral_dixon_coles_WC.py:83-232 D = 13 + 2K + 6T + Cf */
  BPLX_DYNAMIC = 4      /* bpl/dynamic_dixon_coles.py:63-247  D = 10G + 2 + 2K + 7GT */
} bplx_model;

/* memory layout of theta / grad */
typedef enum {
  BPLX_CHAIN_MAJOR = 0, /* theta[c * D + d]   -- what a vmapped jax.ffi call hands over ([C, D]) */
  BPLX_CHAIN_MINOR = 1  /* theta[d * ld + c]  -- native, fully coalesced ([D, ld], ld >= C)      */
} bplx_layout;

#define BPLX_FLAG_DYNAMIC_AS_WRITTEN 1u /* reproduce dynamic_dixon_coles.py:192-218 literally */

/*
 * Static match data = the positional arguments of the reference `_model` after `fit`'s host prep
 * (parse_teams, bpl/_util.py:115-135).  All arrays are HOST arrays of length num_matches unless
 * stated; the library copies what it needs, the caller keeps ownership.
 */
typedef struct {
  int32_t model;            /* bplx_model */
  int32_t num_matches;      /* M */
  int32_t num_teams;        /* T */
  int32_t num_covariates;   /* K (0 = no team_covariates) */
  int32_t num_conferences;  /* Cf (NEUTRAL_WC only, else 0) */
  int32_t num_gameweeks;    /* G (DYNAMIC only, else 0) */
  uint32_t flags;
  const uint16_t* home_team;      /* [M] index into sorted team names */
  const uint16_t* away_team;      /* [M] */
  const uint8_t* home_goals;      /* [M] */
  const uint8_t* away_goals;      /* [M] */
  const uint8_t* neutral_venue;   /* [M] 1 = neutral; NULL = all 0 (required NULL-able for DC/EXT) */
  const uint8_t* home_conf;       /* [M] NEUTRAL_WC, else NULL */
  const uint8_t* away_conf;       /* [M] NEUTRAL_WC, else NULL */
  const int32_t* gameweek;        /* [M] DYNAMIC (0-based, < G), else NULL */
  const float* weights;           /* [M] match weights (time decay x game weights), NULL = unweighted
                                     (the `to_event(1)` branch, extended_dixon_coles.py:216-227) */
  const float* covariates;        /* [T*K] row-major STANDARDISED covariates
                                     (extended_dixon_coles.py:124-127), NULL iff K == 0 */
} bplx_problem_desc;

typedef struct bplx_problem bplx_problem;

/* ---- lifetime ---------------------------------------------------------------------------- */
int bplx_problem_create(const bplx_problem_desc* desc, bplx_problem** out);
void bplx_problem_destroy(bplx_problem* p);

/* number of unconstrained parameters D of the model (numpyro's flattened latent sites) */
int bplx_num_params(const bplx_problem* p);

/*
 * Layout of the flat unconstrained vector, as "name:offset:count:transform;" records
 * (transform in real|exp|sigmoid; names are numpyro's latent site names).  Pointer is owned by
 * the handle.  Replaces numpyro's ravel of the latent-site dict for this path.
 */
const char* bplx_problem_layout(const bplx_problem* p);

/*
 * Plan statistics for benchmarks / DESIGN.md: out[0..7] = matches, phase-1 entries, padded phase-1
 * entries, phase-2 (tau) entries, padded phase-2 entries, shared-memory bytes per CTA, warps per
 * CTA, (team, confederation) pairs.  Returns the number of values written.
 */
int bplx_problem_stats(const bplx_problem* p, long long* out, int n);

/*
 * What each warp of a CTA walks (static models, the plan for one CTA per group of 32 chains): out[w*12 + 0..5] =
 * phase-1 {home-form pieces, away-form pieces, home-form entries, away-form entries, teams, stages}, out[w*12 + 6..11]
 * = phase-2 {pieces, tau = 1 - c X Y entries, tau = 1 + c X entries, tau = 1 + c Y entries, teams, stages} (padded
 * counts).  Returns the number of values written (<= n); 0 for BPLX_DYNAMIC.  For tuning the plan's cost model.
 */
int bplx_problem_warp_stats(const bplx_problem* p, long long* out, int n);

/* ---- log-density + gradient: replaces value_and_grad(potential_fn) per leapfrog ---------- */
/* (numpyro potential_energy of `_model`; the returned lp is the log JOINT density, i.e. MINUS
 * the potential energy, and grad is d lp / d theta.) */
/* Workspace for num_chains chains.  For batches of num_chains x D >= 2^17 elements this includes room for one
 * transposed copy of theta and of the gradient: BPLX_CHAIN_MAJOR calls are then computed in the kernel's native
 * BPLX_CHAIN_MINOR layout (two tiled transposes around the kernel; the kernel alone is 2.2x slower on [chains, D] buffers
 * of configs[2] size).  A smaller workspace (the kernel's own need) is accepted: the call then runs on the buffers as
 * they are. */
size_t bplx_logdensity_workspace_bytes(const bplx_problem* p, int num_chains);

/*
 * Asynchronous on `stream`, capturable in a CUDA graph.  The kernel is launched with programmatic stream
 * serialization: its set-up (parameters, static plan) may overlap the end of the previous kernel of the stream, and
 * everything that reads theta or writes an output waits for that kernel to complete -- stream order is what the caller
 * sees.  Results are a pure function of the inputs (bit-reproducible).  A log-density that overflowed float32 far from
 * the typical set is reported as NaN, never as +inf.  While the call runs, grad doubles as scratch.
 */

int bplx_logdensity_fwdbwd(const bplx_problem* p, int num_chains, int layout, int ld,
                           const float* theta,   /* device, [C, D] or [D, ld] */
                           float* lp,            /* device, [C] */
                           float* grad,          /* device, same layout as theta */
                           float* corr_coef,     /* device, [C] (deterministic site "corr_coef"), may be NULL */
                           void* workspace, size_t workspace_bytes,
                           void* stream);

/*
 * Likelihood-only entry point: what a numpyro `_model` that KEEPS its prior sites hands to `numpyro.factor`
 * (bpl/dixon_coles.py:79-84; it replaces the rates -> Poisson -> tau block, e.g. bpl/neutral_dixon_coles_WC.py:188-232).
 * Input per chain: the CONSTRAINED per-team tables packed as bplx_loglik_layout() says
 *     attack[T] | defence[T] | home_advantage[1 or T]  or  home_attack, away_attack, home_defence, away_defence [T each]
 *     | confederation_strength[Cf] | corr_coef_raw (in (0,1))
 * Output: loglik[C] = sum_m w (Poisson log-pmfs) + sum_m w log tau, corr_coef[C] (the deterministic site), and
 * grad = d loglik / d (every input), same layout as the input -- the cotangent a jax.custom_vjp multiplies through.
 * No priors, no transforms, no Jacobians: those stay numpyro's.  Same launch / workspace rules as
 * bplx_logdensity_fwdbwd (bplx_logdensity_workspace_bytes gives the workspace).  DYNAMIC: BPLX_E_UNSUPPORTED.
 */
int bplx_loglik_num_inputs(const bplx_problem* p);
const char* bplx_loglik_layout(const bplx_problem* p); /* "name:offset:count:real;" records, "unit" for corr_coef_raw */
int bplx_loglik_fwdbwd(const bplx_problem* p, int num_chains, int layout, int ld,
                       const float* tables,  /* device, [C, Dl] or [Dl, ld], Dl = bplx_loglik_num_inputs */
                       float* loglik,        /* device, [C] */
                       float* grad,          /* device, same layout as tables */
                       float* corr_coef,     /* device, [C], may be NULL */
                       void* workspace, size_t workspace_bytes, void* stream);

/* host-buffer variant (chain-major [C, D]); copies in/out inside the call, returns when done.  Page-locked (pinned)
 * arrays are copied by DMA without staging, and lp / corr_coef are then written by the kernel straight into them. */
int bplx_logdensity_fwdbwd_host(bplx_problem* p, int num_chains,
                                const float* theta, float* lp, float* grad, float* corr_coef);

/* ---- posterior-predictive score grid: replaces predict_score_grid_proba/outcome_proba ---- */
/*
 * Posterior sample arrays are [S, T] row-major (the attributes `fit` stores,
 * e.g. neutral_dixon_coles_WC.py:308-334), device pointers; unused ones NULL:
 *   DIXON_COLES : attack, defence, home_advantage [S] (in `home_attack`), corr_coef
 *   EXTENDED    : attack, defence, home_advantage [S,T] (in `home_attack`), corr_coef
 *   NEUTRAL(_WC): attack, defence, home_attack, away_attack, home_defence, away_defence,
 *                 (confederation_strength [S, Cf]), corr_coef
 */
typedef struct {
  int32_t model;           /* bplx_model (DYNAMIC unsupported: reference predict is unusable) */
  int32_t num_samples;     /* S (this rank's shard) */
  int32_t num_teams;       /* T */
  int32_t num_conferences; /* Cf */
  const float* attack;
  const float* defence;
  const float* home_attack;
  const float* away_attack;
  const float* home_defence;
  const float* away_defence;
  const float* confederation_strength;
  const float* corr_coef;  /* [S] */
} bplx_samples;

typedef struct {
  int32_t num_fixtures;          /* F */
  const uint16_t* home_team;     /* device [F] */
  const uint16_t* away_team;     /* device [F] */
  const uint8_t* home_conf;      /* device [F] or NULL */
  const uint8_t* away_conf;      /* device [F] or NULL */
  const uint8_t* neutral_venue;  /* device [F] or NULL */
} bplx_fixtures;

size_t bplx_score_grid_workspace_bytes(const bplx_samples* s, const bplx_fixtures* f, int max_goals);

/*
 * grid[f, hg, ag] = scale * sum_s tau * Pois(hg; lam_h) * Pois(ag; lam_a), 0 <= hg, ag <= max_goals.
 * scale = 1/S for a single-rank call (the reference's mean over samples); a sample-sharded caller
 * passes 1/S_total and all-reduces (sum) the grids.  outcome (optional, [F, 3] = home_win, draw,
 * away_win) is summed from the SAME rank-local grid, so it is also additive across ranks.
 */
int bplx_score_grid(const bplx_samples* s, const bplx_fixtures* f, int max_goals, float scale,
                    float* grid,     /* device [F, g, g], g = max_goals + 1 */
                    float* outcome,  /* device [F, 3] or NULL */
                    void* workspace, size_t workspace_bytes, void* stream);

/*
 * Same, with flags.  BPLX_GRID_REUSE_TABLES: the workspace still holds the per-(sample, team) exponential tables a previous
 * bplx_score_grid[_ex] call built from THESE samples (same arrays, same shapes, same workspace pointer): skip the table
 * pre-pass.  For callers that cut the fixtures into ranges (e.g. to overlap a cross-GPU all-reduce with the next range).
 */
#define BPLX_GRID_REUSE_TABLES 1u
int bplx_score_grid_ex(const bplx_samples* s, const bplx_fixtures* f, int max_goals, float scale, float* grid,
                       float* outcome, void* workspace, size_t workspace_bytes, void* stream, unsigned flags);

/* host-buffer variant: every pointer in s / f / grid / outcome is a HOST pointer. */
int bplx_score_grid_host(const bplx_samples* s, const bplx_fixtures* f, int max_goals, float scale,
                         float* grid, float* outcome);

/* ---- misc --------------------------------------------------------------------------------- */
/* ---- predict over sample shards: the sum of the partial grids over the ranks, over peer memory ------------- */
/*
 * One-shot all-reduce(sum) of two element ranges of SYMMETRIC buffers (one per rank, same size, each mapped into every
 * rank's address space -- e.g. torch.distributed._symmetric_memory): replaces the all_reduce that finishes the reference's
 * `.mean(axis=0)` over posterior samples (bpl/base.py:94-110) when the samples are sharded over GPUs.
 *   bufs[q]     device pointer (valid on THIS rank) to rank q's buffer, q = 0 .. nranks-1; bufs[rank] is this rank's own.
 *               Layout of every buffer: `flag_bytes` of 32-bit flags (zero before the first call; flag p of rank q's
 *               buffer = the last epoch rank p has published), then the float data.
 *   off, cnt    element ranges of the data (grid rows, outcome rows of a fixture range); cnt1 may be 0
 *   out0, out1  this rank's results (ordinary device memory): out[i] = sum over q, in rank order, of data_q[off + i] --
 *               the same bits on every rank
 *   epoch       a counter that every rank increases by one per call (all ranks make the same sequence of calls); 0 = the
 *               kernel keeps the counter itself, in word 16 of this rank's flag area (the call then has no argument that
 *               changes from one call to the next and can be replayed in a CUDA graph).  flag_bytes >= 68, multiple of 16.
 * The kernel publishes this rank's flag to all peers, waits for theirs, then reads every rank's range over NVLink.  A
 * rank may overwrite a range of its own buffer again only after a later call on that rank has returned from its wait
 * (double-buffer by call parity, as bpl_next_b200.parallel.ShardedScoreGrid does).
 */
int bplx_peer_sum(void* const* bufs, int nranks, int rank, size_t flag_bytes, size_t off0, size_t cnt0, float* out0,
                  size_t off1, size_t cnt1, float* out1, unsigned epoch, void* stream);

/* The library reads its testing / tuning switches (BPLX_NO_PDL, BPLX_SPLIT, BPLX_HOST_CHUNKS, BPLX_NUTS_GENERIC,
 * BPLX_NO_TAIL_SPLIT, BPLX_NO_TRANSPOSE) from the environment once, at first use -- never on a launch path; call this after changing them. */
void bplx_reload_env(void);
const char* bplx_last_error(void);
int bplx_version(void);
/* number of kernel launches this library has enqueued from this process (for bench accounting) */
unsigned long long bplx_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BPLX_H_ */
