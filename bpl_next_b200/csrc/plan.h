// plan.h -- the static execution plan of the log-density kernel (K1).
//
// Built once per problem on the host (plan.cc), uploaded by api.cu, consumed by logdensity.cu.
// DESIGN.md "K1" explains the reasoning; the vocabulary in short:
//
//   virtual team v : a (team, confederation) pair seen in the data (just the team when the model
//                    has no confederations).  Per-chain tables are indexed by v.
//   exponent       : one of the six per-virtual-team log-rate halves.  Every match rate factors as
//                    lambda = exp(own exponent) * exp(opponent exponent):
//                      home venue (n=1): lambda_h = exp(A_h1[h]) exp(B_a1[a]),  lambda_a = exp(B_h1[h]) exp(A_a1[a])
//                      neutral    (n=0): lambda_h = exp(A_0[h])  exp(B_0[a]),   lambda_a = exp(B_0[h])  exp(A_0[a])
//                    with A_h1 = att+home_attack+conf, B_h1 = -def-home_defence-conf,
//                         B_a1 = -def-away_defence-conf, A_a1 = att+away_attack+conf,
//                         A_0 = att+conf, B_0 = -def-conf.
//   table          : rows of float2 x 32 chains in shared memory holding the exponentials:
//                      P1[v] = (e^A_h1, e^B_h1)   Q1[v] = (e^B_a1, e^A_a1)   P0[v] = (e^A_0, e^B_0)
//                    each table has one extra all-zero row used by padding entries.
//   list           : the matches of one virtual team in one role at one venue class (kind H1, A1,
//                    H0, A0), cut into pieces of at most kListMax entries.  Every match is in exactly
//                    two lists, so every per-team sum is accumulated in registers by the one warp
//                    that owns the team: no atomics between warps, deterministic.
//                    Inside a list the two rates of an entry are X = own'.x * opp.x and
//                    Y = own'.y * opp.y, where own' is the own row (swapped for the neutral kinds).
//   entry          : (byte offset of the opponent row, weight); identical (list, opponent) pairs
//                    are merged with summed weights.
//   (DYNAMIC: the streams are organised by gameweek instead -- see plan_dynamic.inc.)
//   stream         : what a warp reads in one phase: its lists back to back, fetched stage by stage
//                    (stage_bytes each) through the warp's TMA ring.  A list is stored as pieces, each
//                    a 16-byte ListHdr followed by its entries (a multiple of 16 bytes); a piece never
//                    straddles a stage boundary, so the kernel walks a stage with plain pointer
//                    arithmetic.  Phase 1 = rates (Poisson part + maxima), phase 2 = tau matches.
//   raw slot       : while the phases run, the caller's grad entries of a team's own parameters
//                    (attack/defence pair and the venue effects) hold the gradient with respect to
//                    the team's log-rate halves; the final team pass turns them into parameter
//                    gradients (priors, chain rule) in place.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/bplx.h"

namespace bplx {

constexpr int kChains = 32;       // chains per CTA: one lane per chain
constexpr int kRowBytes = 256;    // one table row: float2 x 32 lanes
constexpr int kMaxCov = 16;       // covariates supported by the kernel
constexpr int kMaxWarps = 24;     // warps per CTA (register budget: 65536 / (24*32) = 85 per thread)
constexpr int kMaxWarpsDyn = 20;  // DYNAMIC kernel: one team (or a few) per warp, 102 registers per thread
constexpr int kStages = 2;        // ring depth per warp (stage size: KernelParams::stage_bytes, 512 or 1024)
constexpr int kAccRows = 13;      // hyper accumulators: lp, mu_d, ls_a, ls_d, mu[4], ls[4], rho
constexpr int kPartRows = 16;     // rows per warp in the final cross-warp reduction
constexpr int kMaxSplit = 8;      // CTAs (one thread-block cluster) that can share one group of 32 chains
constexpr int kNumSplits = 4;     // plans are built for 1, 2, 4 and 8 CTAs per chain group

enum Kind : uint8_t { kH1 = 0, kA1 = 1, kH0 = 2, kA0 = 3 };
enum Exponent : int { eAh1 = 0, eBh1 = 1, eBa1 = 2, eAa1 = 3, eA0 = 4, eB0 = 5 };

constexpr uint8_t kTeamFirst = 1;  // first piece of its team in this warp's stream: clear the accumulators
constexpr uint8_t kTeamLast = 2;   // last one: write the team's raw slots
constexpr uint8_t kVteamLast = 4;  // last piece of its virtual team: write the confederation scratch
constexpr uint8_t kGwFirst = 8;    // DYNAMIC: marker piece (no entries) in front of a warp's pieces of one gameweek; n0 = how many follow
constexpr uint8_t kStageEnd = 16;  // DYNAMIC: filler -- the rest of this stage holds nothing
constexpr uint8_t kPhase2 = 32;    // DYNAMIC (backward stream): tau piece (n0 / n1 / n2 classes); else a rate piece (n0 entries)

// phase-1 entry (plain) and phase-2 entry: 8 bytes
struct Entry {
  uint32_t off;  // byte offset of the opponent row inside the table area
  float w;       // (merged) match weight
};
// phase-1 entry of a clipping model (Extended): 16 bytes
struct EntryClip {
  uint32_t off;
  float w;
  float wyx;  // sum of w * goals of the X rate
  float wyy;  // sum of w * goals of the Y rate
};

struct ListHdr {  // 16 bytes, in front of the entries of every list piece
  uint32_t own_off;  // byte offset of the own row
  uint16_t vteam;  // DYNAMIC: the gameweek
  uint8_t kind;
  uint8_t flags;
  uint16_t n0;  // phase 1: entries (even for 8-byte entries) | phase 2: tau = 1 - c X Y entries (even)
  uint16_t n1;  // phase 2: tau = 1 + c X entries (even); `off` addresses the .x float of the opponent row
  uint16_t n2;  // phase 2: tau = 1 + c Y entries (even); `off` addresses the .y float
  uint16_t team;
};
static_assert(sizeof(ListHdr) == 16, "ListHdr must be 16 bytes");

// offsets of the sites inside the flat unconstrained vector (-1 = absent)
struct ThetaOffsets {
  int mean_defence, log_std_attack, log_std_defence;
  int mean[4];     // DC: home_advantage | EXT: mean_home_advantage | NEU: mean_{ha,aa,hd,ad}
  int log_std[4];  // EXT: std_home_advantage | NEU: std_{ha,aa,hd,ad}
  int u;           // logit u (rho)
  int beta_a, beta_d;  // [K]
  int za, zd;          // [T] (DC: attack_decentered / defence_decentered)
  int dec[4];          // [T] EXT: home_advantage_decentered | NEU: {ha,aa,hd,ad}_decentered
  int conf;            // [Cf]
  int raw;             // logit corr_coef_raw
};

// a scalar hyper-parameter site: Normal(loc, scale) on theta, or HalfNormal(scale) on exp(theta);
// the normalising constants are folded into KernelParams::const_term
struct HyperDesc {
  int off;   // offset in theta
  int row;   // accumulator row holding the likelihood part of its gradient
  int kind;  // 0 = normal, 1 = half-normal on exp(theta) (+ Jacobian), 2 = no prior (likelihood-only entry point)
  float loc, inv_scale;
};

// everything the kernel needs; device pointers are filled by api.cu after upload
// byte ranges of the virtual warps' phase-1 / phase-2 streams for the split in use, as a kernel parameter (constant bank):
// a warp needs them to start its TMA ring, and a global load there is a round trip to L2 on the critical path
struct WarpBounds {
  uint32_t b1[kMaxSplit * kMaxWarps + 1], b2[kMaxSplit * kMaxWarps + 1];
};

struct KernelParams {
  int model, T, K, Cf, V;
  int G;           // DYNAMIC: gameweeks (else 0)
  int as_written;  // DYNAMIC: reproduce dynamic_dixon_coles.py:192-218 literally (attack = defence = 0 in the rates)
  int D, nwarps;
  int ndec;        // decentred per-team venue sites: 0 (DC), 1 (EXT), 4 (NEU, WC)
  int clip;        // Extended: rates clipped at 15
  int lik_only;    // bplx_loglik_fwdbwd: the inputs are the CONSTRAINED per-team tables (attack, defence, venue effects,
                   // confederation strengths, corr_coef_raw in (0,1)); likelihood + tau only, no priors, no Jacobians
  int has1, has0;  // venue classes present
  uint32_t tabP1, tabQ1, tabP0;  // byte offsets of the tables (row V of each = zero row)
  uint32_t tab_bytes;            // table area (reused by the epilogue); DYNAMIC: per-warp table bytes
  uint32_t stage_bytes;          // TMA stage size of the per-warp rings
  uint32_t min_piece1, min_piece2;  // a stage tail shorter than this holds no piece (phase 1 / phase 2)
  uint32_t smem_ring, smem_bar, smem_red, smem_total;  // carve-up (bytes)
  uint32_t epi_team, epi_part;   // epilogue reuse of the table area: team rows, per-warp partials
  uint32_t epi_cl;               // ... and the per-CTA partials of a cluster ([kMaxSplit][kPartRows][32], on rank 0)
  uint32_t smem_red_cl;          // [kMaxSplit][32] d/d corr_coef partials of a cluster (rank 0)
  uint32_t smem_p2;              // 0, or [2 sides][T][2 + ndec][32] f32: phase-2 slot sums when the split-1 plan deals a
                                 // team's home-side and away-side tau lists to different warps (small T, one CTA)
  uint32_t dyn_state, dyn_part, dyn_hyp, dyn_fin;  // DYNAMIC: [T][2][32] f32 walk state, [2][W][10][32] f32 hyper partials,
                                                   // [G][16][32] f32 per-gameweek hyper-parameters, [T][8][32] f32 site stash
  int dyn_nbuf;                           // DYNAMIC: table buffers (2: gameweek j+1 is built while j is still read)
  int dyn_hyp_ws;                         // DYNAMIC: the hyper-parameter table lives in the workspace ([G][16][Cpad] behind the
                                          // prefix sums) because it does not fit in shared memory
  int split_hint;                // largest cluster size worth using (small plans are latency-bound: 1)
  int group0, ngroups;           // chain groups of this launch: [group0, group0 + ngroups) (a call may take two launches)
  int force_clip_forms;          // testing (env BPLX_CLIP_FORMS at create): bit 0 / bit 1 = phase 1 / phase 2 always take the
                                 // clipping form of the arithmetic, even when no rate of the chains is near the clip
  ThetaOffsets off;
  const unsigned char* stream1;  // phase-1 streams of all warps
  const unsigned char* stream2;  // phase-2 streams
  const uint32_t* warp_b1;       // [nwarps+1] byte range of each warp's phase-1 stream
  const uint32_t* warp_b2;       // [nwarps+1]
  const int32_t* team_vptr;      // [T+1] virtual teams of each team (CSR over v, v sorted by team)
  const uint8_t* team_flags;     // [T] bit 0: the team has lists (phase 1 writes its raw slots)
  const uint16_t* v_team;        // [V]
  const uint8_t* v_conf;         // [V]
  const int32_t* conf_vptr;      // [Cf+1]
  const int32_t* conf_vlist;     // [V]
  const float* yexp;             // [V*6] static sums of w*goals per exponent (zero for clip models)
  const float* yteam;            // [T*8] the same folded per team: d/d att, d/d def, d/d venue effect[4], 0, 0
  const float* yconf;            // [Cf]  ... and per confederation
  const float* Xs;               // [T*K]
  const int32_t* gw_tptr;        // DYNAMIC [G+1]: teams with matches in each gameweek (CSR)
  const uint16_t* gw_tlist;      // DYNAMIC
  HyperDesc hyper[12];           // scalar hyper-parameter sites (priors + chain rule in the epilogue)
  int nhyper;
  float w11;         // sum of weights of 1-1 matches
  float const_term;  // -sum w (lgamma(yh+1) + lgamma(ya+1)) + every normalising constant of the priors
  // call arguments
  int C;
  int split;   // CTAs per chain group for this call (cluster size; the streams below belong to this split)
  int sd, sc;  // element strides of theta/grad: index = d*sd + c*sc (api.cu checks that it fits 31 bits)
  const float* theta;
  float* lp;
  float* grad;
  float* corr_coef;
  float* scratch;  // [V][Cpad] per-virtual-team A-B gradient (confederation models)
                   // DYNAMIC: [G*T*2][Cpad] attack / defence of every (gameweek, team) (the random walk's prefix sums)
  int Cpad;
};

struct SplitStreams {  // the two phases' streams for nwarps * split virtual warps
  std::vector<unsigned char> stream1, stream2;
  std::vector<uint32_t> warp_b1, warp_b2;
};

struct HostPlan {
  KernelParams kp{};  // scalar fields filled; pointers null
  // the likelihood-only view of the same plan (static models): input layout [attack T | defence T | venue effects
  // ndec x T (DIXON_COLES: home_advantage, 1) | confederation strengths Cf | corr_coef_raw], see bplx_loglik_fwdbwd
  ThetaOffsets lik_off{};
  HyperDesc lik_hyper[12]{};
  int lik_nhyper = 0, lik_D = 0;
  float lik_const = 0.0f;
  std::string lik_layout;
  std::vector<unsigned char> stream1, stream2;  // split 1
  std::vector<uint32_t> warp_b1, warp_b2;
  SplitStreams more[kNumSplits - 1];            // splits 2, 4, 8 (empty for DYNAMIC)
  std::vector<int32_t> team_vptr, conf_vptr, conf_vlist;
  std::vector<uint8_t> team_flags;
  std::vector<uint16_t> v_team;
  std::vector<uint8_t> v_conf;
  std::vector<float> yexp, yteam, yconf, Xs;
  std::vector<int32_t> gw_tptr;
  std::vector<uint16_t> gw_tlist;
  std::string layout;  // "name:offset:count:transform;" records
  // statistics
  long long n1 = 0, n2 = 0, n1_padded = 0, n2_padded = 0, nlists1 = 0, nlists2 = 0;
};

// Builds the plan; returns BPLX_OK or a negative status with the message in `err`.
// `force_warps` > 0 overrides the warp-count choice (tests, tuning).
int build_plan(const bplx_problem_desc& d, HostPlan* out, std::string* err, int force_warps = 0);

}  // namespace bplx
