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
//                    that owns the virtual team: no atomics, deterministic.
//                    Inside a list the two rates of an entry are X = own'.x * opp.x and
//                    Y = own'.y * opp.y, where own' is the own row (swapped for the neutral kinds).
//   entry          : (byte offset of the opponent row, weight); identical (list, opponent) pairs
//                    are merged with summed weights.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/bplx.h"

#if defined(__CUDACC__)
#define BPLX_HD __host__ __device__
#else
#define BPLX_HD
#endif

namespace bplx {

constexpr int kChains = 32;       // chains per CTA: one lane per chain
constexpr int kRowBytes = 256;    // one table row: float2 x 32 lanes
constexpr int kListMax = 128;     // entries per list piece (bounds the arg-max rescan)
constexpr int kMaxCov = 8;        // covariates supported by the kernel
constexpr int kMaxWarps = 16;
constexpr int kRedRows = 8;       // per-warp rows in the reduction area
constexpr int kStageBytes = 512;  // one TMA bulk copy of a warp's entry stream
constexpr int kStages = 2;        // ring depth per warp

enum Kind : uint8_t { kH1 = 0, kA1 = 1, kH0 = 2, kA0 = 3 };
enum Exponent : int { eAh1 = 0, eBh1 = 1, eBa1 = 2, eAa1 = 3, eA0 = 4, eB0 = 5 };

constexpr uint8_t kListFirst = 1;  // first list of its virtual team in this warp's sequence
constexpr uint8_t kListLast = 2;   // last one: apply the accumulated exponent gradients

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

struct List {  // 32 bytes, read warp-uniformly
  uint32_t ent;      // first entry
  uint32_t n;        // phase 1: number of entries (even; padded with zero-row entries)
  uint32_t own_off;  // byte offset of the own row
  uint32_t vteam;
  uint16_t n_xy, n_x, n_y;  // phase 2: entries per tau class (each even), stored in this order
  uint8_t kind;
  uint8_t flags;
  uint32_t pad[2];
};
static_assert(sizeof(List) == 32, "List must be 32 bytes");

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

// everything the kernel needs; device pointers are filled by api.cu after upload
struct KernelParams {
  int model, T, K, Cf, V;
  int D, nwarps;
  int clip;        // Extended: rates clipped at 15
  int has1, has0;  // venue classes present
  uint32_t tabP1, tabQ1, tabP0;  // byte offsets of the tables (row V of each = zero row)
  uint32_t tab_bytes;
  uint32_t smem_ring, smem_bar, smem_red, smem_total;  // carve-up (bytes)
  ThetaOffsets off;
  const List* lists1;
  const List* lists2;
  const void* ent1;    // Entry or EntryClip
  const Entry* ent2;
  const int32_t* warp_l1;  // [nwarps+1] list range of each warp, phase 1
  const int32_t* warp_l2;  // [nwarps+1]
  const int32_t* warp_e1;  // [nwarps+1] entry range of each warp (contiguous stream), phase 1
  const int32_t* warp_e2;  // [nwarps+1]
  const int32_t* team_vptr;   // [T+1] virtual teams of each team (CSR over v, v sorted by team)
  const uint16_t* v_team;     // [V]
  const uint8_t* v_conf;      // [V]
  const int32_t* conf_vptr;   // [Cf+1]
  const int32_t* conf_vlist;  // [V]
  const float* yexp;          // [V*6] static sums of w*goals per exponent (zero for clip models)
  const float* Xs;            // [T*K]
  float w11;         // sum of weights of 1-1 matches
  float const_term;  // -sum w (lgamma(yh+1) + lgamma(ya+1))
  // call arguments
  int C;
  long long sd, sc;  // element strides of theta/grad: index = d*sd + c*sc
  const float* theta;
  float* lp;
  float* grad;
  float* corr_coef;
  float* scratch;  // [V][Cpad] per-virtual-team A-B gradient (confederation models)
  int Cpad;
};

BPLX_HD inline int hyper_rows(int K) { return 16 + 2 * K; }

struct HostPlan {
  KernelParams kp{};  // scalar fields filled; pointers null
  std::vector<List> lists1, lists2;
  std::vector<Entry> ent1, ent2;
  std::vector<EntryClip> ent1c;
  std::vector<int32_t> warp_l1, warp_l2, warp_e1, warp_e2, team_vptr, conf_vptr, conf_vlist;
  std::vector<uint16_t> v_team;
  std::vector<uint8_t> v_conf;
  std::vector<float> yexp, Xs;
  std::string layout;  // "name:offset:count:transform;" records
  // statistics
  long long n1 = 0, n2 = 0, n1_padded = 0, n2_padded = 0;
};

// Builds the plan; returns BPLX_OK or a negative status with the message in `err`.
int build_plan(const bplx_problem_desc& d, HostPlan* out, std::string* err);

}  // namespace bplx
