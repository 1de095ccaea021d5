// plan.cc -- host-side plan builder for the log-density kernel (see plan.h for the vocabulary).
//
// Replaces, for this path, the trace-time work numpyro/XLA do on the reference `_model`s
// (bpl/dixon_coles.py:39-84, bpl/extended_dixon_coles.py:78-248, bpl/neutral_dixon_coles.py:102-283,
// bpl/neutral_dixon_coles_WC.py:83-232): the static gathers `attack[home_team]` ... and the boolean
// score masks of bpl/_util.py:58-87 become per-team entry lists, and the theta-independent sums
// (sum w*y per team, sum w*lgamma(y+1), weight of the 1-1 matches) are folded on the host in double.
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>

namespace bplx {
namespace {

struct RawEntry {
  uint32_t opp;  // opponent virtual team
  uint8_t cls;   // phase 2: 0 = XY, 1 = X, 2 = Y
  double w, wyx, wyy;
};

std::string fmt(const char* f, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof buf, f, ap);
  va_end(ap);
  return buf;
}

void add_site(std::string* s, const char* name, int off, int count, const char* tr) {
  *s += fmt("%s:%d:%d:%s;", name, off, count, tr);
}

// theta layout == oracle/models.py:site_layout (declaration order of each reference `_model`)
int make_layout(int model, int T, int K, int Cf, ThetaOffsets* o, std::string* s) {
  ThetaOffsets z;
  memset(&z, 0xff, sizeof z);  // all -1
  int d = 0;
  auto scalar = [&](const char* n, const char* tr) { add_site(s, n, d, 1, tr); return d++; };
  auto vec = [&](const char* n, int cnt, const char* tr) {
    add_site(s, n, d, cnt, tr);
    int r = d;
    d += cnt;
    return r;
  };
  if (model == BPLX_DIXON_COLES) {  // dixon_coles.py:46-78
    z.mean[0] = scalar("home_advantage", "real");
    z.mean_defence = scalar("mean_defence", "real");
    z.log_std_attack = scalar("std_attack", "exp");
    z.log_std_defence = scalar("std_defence", "exp");
    z.za = vec("attack_decentered", T, "real");
    z.zd = vec("defence_decentered", T, "real");
    z.raw = scalar("corr_coef_raw", "sigmoid");
  } else if (model == BPLX_EXTENDED) {  // extended_dixon_coles.py:112-235
    z.mean[0] = scalar("mean_home_advantage", "real");
    z.log_std[0] = scalar("std_home_advantage", "exp");
    z.mean_defence = scalar("mean_defence", "real");
    z.log_std_attack = scalar("std_attack", "exp");
    z.log_std_defence = scalar("std_defence", "exp");
    z.beta_a = vec("attack_coefficients", K, "real");
    z.beta_d = vec("defence_coefficients", K, "real");
    z.u = scalar("u", "sigmoid");
    z.za = vec("standardised_attack", T, "real");
    z.zd = vec("standardised_defence", T, "real");
    z.dec[0] = vec("home_advantage_decentered", T, "real");
    z.raw = scalar("corr_coef_raw", "sigmoid");
  } else if (model == BPLX_NEUTRAL || model == BPLX_NEUTRAL_WC) {  // neutral...:136-272 / ..._WC.py:99-219
    static const char* nm[4] = {"home_attack", "away_attack", "home_defence", "away_defence"};
    z.mean_defence = scalar("mean_defence", "real");
    z.log_std_attack = scalar("std_attack", "exp");
    z.log_std_defence = scalar("std_defence", "exp");
    for (int i = 0; i < 4; i++) z.mean[i] = scalar((std::string("mean_") + nm[i]).c_str(), "real");
    for (int i = 0; i < 4; i++) z.log_std[i] = scalar((std::string("std_") + nm[i]).c_str(), "exp");
    z.u = scalar("u", "sigmoid");
    z.beta_a = vec("attack_coefficients", K, "real");
    z.beta_d = vec("defence_coefficients", K, "real");
    z.za = vec("standardised_attack", T, "real");
    z.zd = vec("standardised_defence", T, "real");
    for (int i = 0; i < 4; i++) z.dec[i] = vec((std::string(nm[i]) + "_decentered").c_str(), T, "real");
    if (model == BPLX_NEUTRAL_WC) z.conf = vec("confederation_strength_decentered", Cf, "real");
    z.raw = scalar("corr_coef_raw", "sigmoid");
  } else if (model == BPLX_DYNAMIC) {  // dynamic_dixon_coles.py:74-241; T here is G * T for the per-(gameweek, team) sites
    static const char* nm[4] = {"home_attack", "away_attack", "home_defence", "away_defence"};
    const int G = Cf;  // (the caller passes G in the Cf slot)
    for (int i = 0; i < 4; i++) z.mean[i] = vec((std::string("mean_") + nm[i]).c_str(), G, "real");
    for (int i = 0; i < 4; i++) z.log_std[i] = vec((std::string("std_") + nm[i]).c_str(), G, "exp");
    z.log_std_attack = vec("std_attack", G, "exp");
    z.log_std_defence = vec("std_defence", G, "exp");
    z.mean_defence = scalar("mean_defence", "real");
    z.beta_a = vec("attack_coefficients", K, "real");
    z.beta_d = vec("defence_coefficients", K, "real");
    z.u = vec("u", G * T, "sigmoid");
    z.za = vec("standardised_attack", G * T, "real");
    z.zd = vec("standardised_defence", G * T, "real");
    for (int i = 0; i < 4; i++) z.dec[i] = vec((std::string(nm[i]) + "_decentered").c_str(), G * T, "real");
    z.raw = scalar("corr_coef_raw", "sigmoid");
  } else {
    return -1;
  }
  if (K == 0) z.beta_a = z.beta_d = -1;
  *o = z;
  return d;
}

}  // namespace

static int build_plan_dynamic(const bplx_problem_desc& d, HostPlan* out, std::string* err, int force_warps);

int build_plan(const bplx_problem_desc& d, HostPlan* out, std::string* err, int force_warps) {
  const int model = d.model, M = d.num_matches, T = d.num_teams, K = d.num_covariates;
  const bool wc = model == BPLX_NEUTRAL_WC;
  const bool neutral = wc || model == BPLX_NEUTRAL;
  const int Cf = wc ? d.num_conferences : 0;
#define FAIL(code, ...)      \
  do {                       \
    *err = fmt(__VA_ARGS__); \
    return (code);           \
  } while (0)
  if (model < 0 || model > BPLX_DYNAMIC) FAIL(BPLX_E_INVALID, "unknown model %d", model);
  if (model == BPLX_DYNAMIC) return build_plan_dynamic(d, out, err, force_warps);
  if (M <= 0) FAIL(BPLX_E_INVALID, "num_matches must be positive (got %d)", M);
  if (T <= 0 || T > 65535) FAIL(BPLX_E_INVALID, "num_teams out of range (%d)", T);
  if (K < 0 || K > kMaxCov) FAIL(BPLX_E_UNSUPPORTED, "num_covariates %d outside [0, %d]", K, kMaxCov);
  if (K > 0 && model == BPLX_DIXON_COLES) FAIL(BPLX_E_INVALID, "DIXON_COLES takes no covariates");
  if (K > 0 && !d.covariates) FAIL(BPLX_E_INVALID, "covariates is NULL but num_covariates = %d", K);
  if (!d.home_team || !d.away_team || !d.home_goals || !d.away_goals)
    FAIL(BPLX_E_INVALID, "home_team / away_team / home_goals / away_goals must not be NULL");
  if (wc && (Cf <= 0 || Cf > 255 || !d.home_conf || !d.away_conf))
    FAIL(BPLX_E_INVALID, "NEUTRAL_WC needs home_conf, away_conf and 0 < num_conferences <= 255");
  for (int m = 0; m < M; m++) {
    if (d.home_team[m] >= T || d.away_team[m] >= T)
      FAIL(BPLX_E_INVALID, "team index out of range at match %d", m);
    if (wc && (d.home_conf[m] >= Cf || d.away_conf[m] >= Cf))
      FAIL(BPLX_E_INVALID, "confederation index out of range at match %d", m);
    if (d.weights && !(d.weights[m] >= 0.0f && d.weights[m] <= 3.0e38f))
      FAIL(BPLX_E_INVALID, "weights must be finite and non-negative (match %d)", m);
  }

  HostPlan& P = *out;
  KernelParams& kp = P.kp;
  kp.model = model;
  kp.T = T;
  kp.K = K;
  kp.Cf = Cf;
  kp.clip = model == BPLX_EXTENDED;
  kp.ndec = model == BPLX_DIXON_COLES ? 0 : (model == BPLX_EXTENDED ? 1 : 4);
  kp.D = make_layout(model, T, K, Cf, &kp.off, &P.layout);

  // ---- scalar hyper-parameter sites and every theta-independent constant --------
  const double kLogSqrt2Pi = 0.91893853320467274178, kLog2 = 0.69314718055994530942;
  double const_term = 0.0;
  double const_lik = 0.0;  // the likelihood's own constant: -sum w (lgamma(yh+1) + lgamma(ya+1))
  {
    int n = 0;
    auto normal = [&](int off, int row, double loc, double scale) {
      kp.hyper[n++] = HyperDesc{off, row, 0, (float)loc, (float)(1.0 / scale)};
      const_term += -std::log(scale) - kLogSqrt2Pi;
    };
    auto halfnormal = [&](int off, int row, double scale) {
      kp.hyper[n++] = HyperDesc{off, row, 1, 0.0f, (float)(1.0 / scale)};
      const_term += -std::log(scale) - kLogSqrt2Pi + kLog2;
    };
    const double std_scale = neutral ? 0.5 : 1.0;  // neutral_dixon_coles.py:138-139
    normal(kp.off.mean_defence, 1, 0.0, 1.0);
    halfnormal(kp.off.log_std_attack, 2, std_scale);
    halfnormal(kp.off.log_std_defence, 3, std_scale);
    if (!neutral) {
      normal(kp.off.mean[0], 4, 0.1, 0.2);
      if (model == BPLX_EXTENDED) halfnormal(kp.off.log_std[0], 8, 1.0);
    } else {
      for (int i = 0; i < 4; i++) {
        normal(kp.off.mean[i], 4 + i, (i & 1) ? -0.1 : 0.1, 0.2);
        halfnormal(kp.off.log_std[i], 8 + i, 1.0);
      }
    }
    kp.nhyper = n;
    const_term -= (double)T * (2.0 + kp.ndec) * kLogSqrt2Pi;   // standardised pair + decentred sites, N(.,1)
    const_term -= (double)(Cf + 2 * K) * kLogSqrt2Pi;          // confederation strengths, covariate coefficients
    if (model != BPLX_DIXON_COLES) const_term += std::log(20.0);  // u ~ Beta(2, 4)
    const_term += std::log(6.0);                                   // corr_coef_raw ~ Beta(2, 2)
  }

  // ---- virtual teams -------------------------------------------------------------------------
  std::map<std::pair<int, int>, int> vmap;  // (team, conf) -> v, ordered by team then conf
  if (wc) {
    for (int m = 0; m < M; m++) {
      vmap[{d.home_team[m], d.home_conf[m]}] = 0;
      vmap[{d.away_team[m], d.away_conf[m]}] = 0;
    }
  } else {
    for (int t = 0; t < T; t++) vmap[{t, 0}] = 0;
  }
  int V = 0;
  P.v_team.clear();
  P.v_conf.clear();
  for (auto& kv : vmap) {
    kv.second = V++;
    P.v_team.push_back((uint16_t)kv.first.first);
    P.v_conf.push_back((uint8_t)kv.first.second);
  }
  if (V > 65535) FAIL(BPLX_E_UNSUPPORTED, "too many (team, confederation) pairs (%d)", V);
  kp.V = V;
  P.team_vptr.assign(T + 1, 0);
  for (int v = 0; v < V; v++) P.team_vptr[P.v_team[v] + 1]++;
  for (int t = 0; t < T; t++) P.team_vptr[t + 1] += P.team_vptr[t];
  P.conf_vptr.assign(Cf + 1, 0);
  P.conf_vlist.assign(wc ? V : 0, 0);
  if (wc) {
    for (int v = 0; v < V; v++) P.conf_vptr[P.v_conf[v] + 1]++;
    for (int c = 0; c < Cf; c++) P.conf_vptr[c + 1] += P.conf_vptr[c];
    std::vector<int> fill(P.conf_vptr.begin(), P.conf_vptr.end() - 1);
    for (int v = 0; v < V; v++) P.conf_vlist[fill[P.v_conf[v]]++] = v;
  }

  // ---- venue classes and table layout ----------------------------------------------------------
  bool has1 = false, has0 = false;
  for (int m = 0; m < M; m++) {
    bool nv = neutral && d.neutral_venue && d.neutral_venue[m];
    (nv ? has0 : has1) = true;
  }
  kp.has1 = has1;
  kp.has0 = has0;
  const uint32_t tab_rows = (uint32_t)V + 1;
  uint32_t off = 0;
  kp.tabP1 = kp.tabQ1 = kp.tabP0 = 0;
  if (has1) {
    kp.tabP1 = off;
    off += tab_rows * kRowBytes;
    kp.tabQ1 = off;
    off += tab_rows * kRowBytes;
  }
  if (has0) {
    kp.tabP0 = off;
    off += tab_rows * kRowBytes;
  }
  const uint32_t table_bytes = off;

  // ---- per-match entries -----------------------------------------------------------------------
  // raw1[v][kind], raw2[v][kind]
  std::vector<std::vector<RawEntry>> raw1((size_t)V * 4), raw2((size_t)V * 4);
  std::vector<double> yexp((size_t)V * 6, 0.0);
  double w11 = 0.0;
  for (int m = 0; m < M; m++) {
    const int h = d.home_team[m], a = d.away_team[m];
    const int hc = wc ? d.home_conf[m] : 0, ac = wc ? d.away_conf[m] : 0;
    const int hv = vmap[{h, hc}], av = vmap[{a, ac}];
    const bool nv = neutral && d.neutral_venue && d.neutral_venue[m];
    const double w = d.weights ? (double)d.weights[m] : 1.0;
    const int yh = d.home_goals[m], ya = d.away_goals[m];
    const_term -= w * (std::lgamma(yh + 1.0) + std::lgamma(ya + 1.0));
    const_lik -= w * (std::lgamma(yh + 1.0) + std::lgamma(ya + 1.0));
    const int kh = nv ? kH0 : kH1, ka = nv ? kA0 : kA1;
    // X/Y meaning per kind: H1, A1, A0: X = lambda_h, Y = lambda_a.  H0: X = lambda_a, Y = lambda_h.
    RawEntry eh{(uint32_t)av, 0, w, nv ? w * ya : w * yh, nv ? w * yh : w * ya};
    RawEntry ea{(uint32_t)hv, 0, w, w * yh, w * ya};
    raw1[(size_t)hv * 4 + kh].push_back(eh);
    raw1[(size_t)av * 4 + ka].push_back(ea);
    if (!kp.clip) {  // static y * log(lambda) part: linear in the exponents
      if (!nv) {
        yexp[(size_t)hv * 6 + eAh1] += w * yh;
        yexp[(size_t)hv * 6 + eBh1] += w * ya;
        yexp[(size_t)av * 6 + eBa1] += w * yh;
        yexp[(size_t)av * 6 + eAa1] += w * ya;
      } else {
        yexp[(size_t)hv * 6 + eA0] += w * yh;
        yexp[(size_t)hv * 6 + eB0] += w * ya;
        yexp[(size_t)av * 6 + eB0] += w * yh;
        yexp[(size_t)av * 6 + eA0] += w * ya;
      }
    }
    // tau classes (bpl/_util.py:58-91)
    if (yh == 1 && ya == 1) {
      w11 += w;
    } else if (yh <= 1 && ya <= 1) {
      // 0-0: 1 - c lh la (XY) ; 1-0: 1 + c la ; 0-1: 1 + c lh
      int cls_std = (yh == 0 && ya == 0) ? 0 : (yh == 1 ? 2 /* needs lambda_a = Y */ : 1 /* lambda_h = X */);
      int cls_h = cls_std;
      if (nv && cls_std != 0) cls_h = 3 - cls_std;  // H0 exchanges X and Y
      eh.cls = (uint8_t)cls_h;
      ea.cls = (uint8_t)cls_std;
      raw2[(size_t)hv * 4 + kh].push_back(eh);
      raw2[(size_t)av * 4 + ka].push_back(ea);
    }
  }
  kp.w11 = (float)w11;
  kp.const_term = (float)const_term;
  {  // likelihood-only view (bplx_loglik_fwdbwd): the sites the reference's rates read, as named by its `fit` attributes
    ThetaOffsets z;
    memset(&z, 0xff, sizeof z);
    int dd = 0;
    std::string& ls = P.lik_layout;
    auto vec = [&](const char* n, int cnt) {
      add_site(&ls, n, dd, cnt, "real");
      const int r0 = dd;
      dd += cnt;
      return r0;
    };
    z.za = vec("attack", T);
    z.zd = vec("defence", T);
    int nh = 0;
    if (model == BPLX_DIXON_COLES) {
      z.mean[0] = vec("home_advantage", 1);
      P.lik_hyper[nh++] = HyperDesc{z.mean[0], 4, 2, 0.0f, 1.0f};
    } else if (model == BPLX_EXTENDED) {
      z.dec[0] = vec("home_advantage", T);
    } else {
      static const char* nm[4] = {"home_attack", "away_attack", "home_defence", "away_defence"};
      for (int i = 0; i < 4; i++) z.dec[i] = vec(nm[i], T);
    }
    if (wc) z.conf = vec("confederation_strength", Cf);
    add_site(&ls, "corr_coef_raw", dd, 1, "unit");
    z.raw = dd++;
    P.lik_off = z;
    P.lik_nhyper = nh;
    P.lik_D = dd;
    P.lik_const = (float)const_lik;
  }
  P.yexp.resize(yexp.size());
  for (size_t i = 0; i < yexp.size(); i++) P.yexp[i] = (float)yexp[i];
  {  // the same static sums folded per team / per confederation
    std::vector<double> yt((size_t)T * 8, 0.0), yc((size_t)Cf, 0.0);
    for (int v = 0; v < V; v++) {
      const double* g = &yexp[(size_t)v * 6];
      double* o = &yt[(size_t)P.v_team[v] * 8];
      const double gA = g[eAh1] + g[eAa1] + g[eA0], gB = g[eBh1] + g[eBa1] + g[eB0];
      o[0] += gA;
      o[1] -= gB;
      o[2] += g[eAh1];
      o[3] += g[eAa1];
      o[4] -= g[eBh1];
      o[5] -= g[eBa1];
      if (wc) yc[P.v_conf[v]] += gA - gB;
    }
    P.yteam.resize(yt.size());
    for (size_t i = 0; i < yt.size(); i++) P.yteam[i] = (float)yt[i];
    P.yconf.resize(yc.size());
    for (size_t i = 0; i < yc.size(); i++) P.yconf[i] = (float)yc[i];
  }
  P.Xs.assign(d.covariates ? d.covariates : nullptr, d.covariates ? d.covariates + (size_t)T * K : nullptr);

  auto own_off = [&](int v, int kind) -> uint32_t {
    uint32_t base = kind == kH1 ? kp.tabP1 : kind == kA1 ? kp.tabQ1 : kp.tabP0;
    return base + (uint32_t)v * kRowBytes;
  };
  auto opp_base = [&](int kind) -> uint32_t { return kind == kH1 ? kp.tabQ1 : kind == kA1 ? kp.tabP1 : kp.tabP0; };

  // merge identical (opponent[, class]) entries of a list
  auto merge = [](std::vector<RawEntry>& v) {
    std::stable_sort(v.begin(), v.end(), [](const RawEntry& a, const RawEntry& b) {
      return a.cls != b.cls ? a.cls < b.cls : a.opp < b.opp;
    });
    size_t o = 0;
    for (size_t i = 0; i < v.size(); i++) {
      if (o && v[o - 1].cls == v[i].cls && v[o - 1].opp == v[i].opp) {
        v[o - 1].w += v[i].w;
        v[o - 1].wyx += v[i].wyx;
        v[o - 1].wyy += v[i].wyy;
      } else {
        v[o++] = v[i];
      }
    }
    v.resize(o);
  };
  for (auto& v : raw1) merge(v);
  for (auto& v : raw2) merge(v);
  P.team_flags.assign(T, 0);
  for (int v = 0; v < V; v++)
    for (int k = 0; k < 4; k++)
      if (!raw1[(size_t)v * 4 + k].empty()) P.team_flags[P.v_team[v]] |= 1;

  // ---- assign teams to warps (longest processing time first) ------------------------------------
  // All lists of a team go to the same warp: the owner thread of (team, chain) is the only one that
  // touches that team's raw slots inside a phase.
  auto team_costs = [&](const std::vector<std::vector<RawEntry>>& raw, double per_entry, double per_list,
                        double per_team) {
    std::vector<double> cost(T, 0.0);
    for (int v = 0; v < V; v++)
      for (int k = 0; k < 4; k++) {
        const size_t n = raw[(size_t)v * 4 + k].size();
        if (n) cost[P.v_team[v]] += per_list * (1.0 + (double)n / 100.0) + per_entry * (double)n;
      }
    for (int t = 0; t < T; t++)
      if (cost[t] > 0.0) cost[t] += per_team;
    return cost;
  };
  auto assign = [&](const std::vector<double>& cost, int W, std::vector<std::vector<int>>* teams) {
    std::vector<int> order(T);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    std::vector<double> load(W, 0.0);
    teams->assign(W, {});
    for (int t : order) {
      if (cost[t] == 0.0) continue;
      int best = 0;
      for (int w = 1; w < W; w++)
        if (load[w] < load[best]) best = w;
      load[best] += cost[t];
      (*teams)[best].push_back(t);
    }
    for (auto& v : *teams) std::sort(v.begin(), v.end());
    return *std::max_element(load.begin(), load.end());
  };
  // rough instruction counts per entry / list piece / team of the two phases
  const std::vector<double> cost1 = team_costs(raw1, kp.clip ? 16.0 : 6.0, 40.0, 30.0);
  const std::vector<double> cost2 = team_costs(raw2, 12.0, 60.0, 30.0);
  // shared memory left for the rings decides how many warps / which stage size fit
  auto smem_need = [&](int W, uint32_t stage) {
    const uint32_t epi = (K > 0 ? (uint32_t)T * kRowBytes : 0u) + (uint32_t)(W + kMaxSplit) * kPartRows * 128u;
    const uint32_t tabb = (std::max(table_bytes, epi) + 127u) / 128u * 128u;
    return tabb + (uint32_t)W * kStages * stage + 128u + (uint32_t)W * kStages * 8u + 3u * 256u + 2u * 128u + 2u * 128u +
           12u * 128u + (uint32_t)(W + kMaxSplit) * 128u;
  };
  const uint32_t kSmemMax = 227u * 1024u;
  int W = 0;
  uint32_t stage = 1024;
  std::vector<std::vector<int>> w1, w2;
  if (force_warps <= 0) {
    if (const char* e = getenv("BPLX_NWARPS")) force_warps = atoi(e);
  }
  if (const char* e = getenv("BPLX_STAGE")) stage = (uint32_t)atoi(e) == 512u ? 512u : 1024u;
  if (force_warps > 0) {
    W = std::max(1, std::min(force_warps, kMaxWarps));
    if (smem_need(W, stage) > kSmemMax) stage = 512;
  } else {
    // the makespan of the slower warp decides; more warps than needed only cost shared memory
    double best = 0.0;
    for (int cand : {4, 8, 12, 16, 20, 24}) {
      if (cand > kMaxWarps) continue;
      uint32_t st = stage;
      if (smem_need(cand, st) > kSmemMax) st = 512;
      if (smem_need(cand, st) > kSmemMax) continue;
      std::vector<std::vector<int>> t1, t2;
      const double span = assign(cost1, cand, &t1) + assign(cost2, cand, &t2) + 4.0 * cand + (st == 512 ? 0.05 * best : 0.0);
      if (W == 0 || span < 0.97 * best) best = span, W = cand, stage = st;
    }
    if (W == 0) W = 4, stage = 512;
  }
  kp.nwarps = W;
  kp.stage_bytes = stage;
  kp.force_clip_forms = 0;
  if (const char* e = getenv("BPLX_CLIP_FORMS")) kp.force_clip_forms = atoi(e);
  {  // a cluster only pays when one CTA's share of the walk is long compared with the cluster barriers (~1 us each):
     // measured, configs[1] data (1.5 k cost units per warp) loses 50 %, configs[2] data (50 k) gains 3.4x
    std::vector<std::vector<int>> t1, t2;
    const double span = assign(cost1, W, &t1) + assign(cost2, W, &t2);
    kp.split_hint = span >= 32000.0 ? 8 : span >= 16000.0 ? 4 : span >= 8000.0 ? 2 : 1;
  }

  // ---- few teams: phase 2 by (team, side) ------------------------------------------------------------
  // With about as many teams as warps the longest tau list decides phase 2 (configs[1]: 13..74 entries per team on 20
  // warps).  The split-1 plan may then deal a team's home-side (H1, H0) and away-side (A1, A0) lists to different
  // warps; each side's slot sums go to its own shared-memory table (plain stores) and the team pass adds the two in a
  // fixed order, so the result stays deterministic.  Not with confederations (their per-virtual-team sums are
  // red.add'ed by the list owner), not in a cluster (the tables would be remote).
  std::vector<std::vector<int>> w2_items;  // per warp: team * 4 + side mask (1 home side, 2 away side, 3 both)
  kp.smem_p2 = 0;
  {
    std::vector<std::pair<int, double>> units;
    for (int t = 0; t < T; t++)
      for (int side = 0; side < 2; side++) {
        double c = 0.0;
        for (int v = P.team_vptr[t]; v < P.team_vptr[t + 1]; v++)
          for (int k = side; k < 4; k += 2) {
            const size_t n = raw2[(size_t)v * 4 + k].size();
            if (n) c += 60.0 * (1.0 + (double)n / 100.0) + 12.0 * (double)n;
          }
        if (c > 0.0) units.push_back({t * 4 + (1 << side), c + 30.0});
      }
    std::stable_sort(units.begin(), units.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
    std::vector<double> load(W, 0.0);
    std::vector<std::vector<int>> items(W);
    for (const auto& u : units) {
      int best = 0;
      for (int w = 1; w < W; w++)
        if (load[w] < load[best]) best = w;
      load[best] += u.second;
      items[best].push_back(u.first);
    }
    for (auto& v : items) std::sort(v.begin(), v.end());
    std::vector<std::vector<int>> t2;
    const double span_team = assign(cost2, W, &t2);
    const double span_unit = units.empty() ? 0.0 : *std::max_element(load.begin(), load.end());
    const uint32_t p2_bytes = 2u * (uint32_t)T * (uint32_t)(2 + kp.ndec) * 128u;
    const bool off = getenv("BPLX_NO_P2SPLIT") != nullptr;
    if (!off && Cf <= 0 && span_unit < 0.85 * span_team && smem_need(W, stage) + p2_bytes <= kSmemMax) {
      w2_items = items;
      kp.smem_p2 = 1;  // (the offset is set in the carve-up below)
    }
  }

  // ---- emit the streams ---------------------------------------------------------------------------
  const uint32_t zero_row = (uint32_t)V * kRowBytes;
  auto put = [](std::vector<unsigned char>* s, const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    s->insert(s->end(), b, b + n);
  };
  // one set of streams per split: Wtot = W * split virtual warps (virtual warp = cluster rank * W + warp)
  auto emit = [&](int Wtot, std::vector<unsigned char>* s1, std::vector<unsigned char>* s2, std::vector<uint32_t>* wb1,
                  std::vector<uint32_t>* wb2) {
  const bool by_side = kp.smem_p2 != 0 && Wtot == kp.nwarps;  // phase 2 of the split-1 plan dealt by (team, side)
  const int W = Wtot;  // (shadows the per-CTA warp count inside the emitter)
  assign(cost1, W, &w1);
  assign(cost2, W, &w2);
  std::vector<std::vector<int>> it1(W), it2(W);
  for (int w = 0; w < W; w++) {
    for (int t : w1[w]) it1[w].push_back(t * 4 + 3);
    for (int t : w2[w]) it2[w].push_back(t * 4 + 3);
  }
  if (by_side) it2 = w2_items;
  P.n1 = P.n2 = P.n1_padded = P.n2_padded = P.nlists1 = P.nlists2 = 0;
  wb1->assign(W + 1, 0);
  wb2->assign(W + 1, 0);
  for (int phase = 1; phase <= 2; phase++) {
    const auto& raw = phase == 1 ? raw1 : raw2;
    const auto& by_warp = phase == 1 ? it1 : it2;
    std::vector<unsigned char>& S = phase == 1 ? *s1 : *s2;
    std::vector<uint32_t>& wb = phase == 1 ? *wb1 : *wb2;
    const size_t esz = (phase == 1 && kp.clip) ? sizeof(EntryClip) : sizeof(Entry);
    const size_t gran = 16;  // bytes of entries a piece grows by
    const size_t min_piece = sizeof(ListHdr) + gran;
    (phase == 1 ? kp.min_piece1 : kp.min_piece2) = (uint32_t)min_piece;
    S.clear();
    for (int w = 0; w < W; w++) {
      const size_t wstart = S.size();
      for (int item : by_warp[w]) {
        const int t = item >> 2, sides = item & 3;
        size_t team_first_hdr = (size_t)-1, last_hdr = 0;
        for (int v = P.team_vptr[t]; v < P.team_vptr[t + 1]; v++) {
          size_t vteam_last_hdr = (size_t)-1;
          for (int k = 0; k < 4; k++) {
            const auto& r = raw[(size_t)v * 4 + k];
            if (r.empty() || !((sides >> (k & 1)) & 1)) continue;
            (phase == 1 ? P.n1 : P.n2) += (long long)r.size();
            // the padded entry sequence of the list, with the tau class of every entry
            std::vector<unsigned char> body;
            std::vector<uint8_t> cls;
            if (phase == 1) {
              for (const RawEntry& e : r) {
                const uint32_t o = opp_base(k) + e.opp * kRowBytes;
                if (kp.clip) {
                  EntryClip x{o, (float)e.w, (float)e.wyx, (float)e.wyy};
                  put(&body, &x, sizeof x);
                } else {
                  Entry x{o, (float)e.w};
                  put(&body, &x, sizeof x);
                }
                cls.push_back(0);
              }
              while (!kp.clip && (cls.size() & 1)) {  // pad to an even count with a zero-row entry
                Entry x{opp_base(k) + zero_row, 0.0f};
                put(&body, &x, sizeof x);
                cls.push_back(0);
              }
            } else {
              size_t i = 0;
              for (int c = 0; c < 3; c++) {
                const uint32_t comp = c == 2 ? 4u : 0u;  // the Y class reads the .y float of the opponent row
                size_t cnt = 0;
                for (; i < r.size() && r[i].cls == c; i++, cnt++) {
                  Entry x{opp_base(k) + r[i].opp * kRowBytes + comp, (float)r[i].w};
                  put(&body, &x, sizeof x);
                  cls.push_back((uint8_t)c);
                }
                if (cnt & 1) {
                  Entry x{opp_base(k) + zero_row + comp, 0.0f};
                  put(&body, &x, sizeof x);
                  cls.push_back((uint8_t)c);
                }
              }
            }
            (phase == 1 ? P.n1_padded : P.n2_padded) += (long long)cls.size();
            // cut into pieces that stay inside a stage; a stage tail shorter than the smallest piece is dead
            size_t done = 0;  // entries emitted
            while (done < cls.size()) {
              size_t in_stage = (S.size() - wstart) % stage;
              if (stage - in_stage < min_piece) {
                S.resize(S.size() + (stage - in_stage), 0);
                in_stage = 0;
              }
              const size_t space = stage - in_stage - sizeof(ListHdr);
              const size_t take = std::min(cls.size() - done, space / gran * gran / esz);
              ListHdr H{};
              H.own_off = own_off(v, k);
              H.vteam = (uint16_t)v;
              H.kind = (uint8_t)k;
              H.team = (uint16_t)t;
              uint16_t cnt[3] = {0, 0, 0};
              for (size_t i = done; i < done + take; i++) cnt[cls[i]]++;
              H.n0 = cnt[0];
              H.n1 = cnt[1];
              H.n2 = cnt[2];
              const size_t hdr_at = S.size();
              put(&S, &H, sizeof H);
              put(&S, body.data() + done * esz, take * esz);
              done += take;
              (phase == 1 ? P.nlists1 : P.nlists2)++;
              if (team_first_hdr == (size_t)-1) team_first_hdr = hdr_at;
              vteam_last_hdr = last_hdr = hdr_at;
            }
          }
          if (vteam_last_hdr != (size_t)-1) reinterpret_cast<ListHdr*>(&S[vteam_last_hdr])->flags |= kVteamLast;
        }
        if (team_first_hdr != (size_t)-1) {
          reinterpret_cast<ListHdr*>(&S[team_first_hdr])->flags |= kTeamFirst;
          reinterpret_cast<ListHdr*>(&S[last_hdr])->flags |= kTeamLast;
        }
      }
      wb[w + 1] = (uint32_t)S.size();
    }
    if (S.empty()) S.resize(16, 0);  // never upload an empty buffer
  }
  };
  for (int i = kNumSplits - 1; i >= 1; i--)
    emit(W << i, &P.more[i - 1].stream1, &P.more[i - 1].stream2, &P.more[i - 1].warp_b1, &P.more[i - 1].warp_b2);
  emit(W, &P.stream1, &P.stream2, &P.warp_b1, &P.warp_b2);  // last: the statistics describe split 1

  // ---- shared-memory carve-up --------------------------------------------------------------------
  kp.epi_team = 0;
  kp.epi_part = K > 0 ? (uint32_t)T * kRowBytes : 0u;
  kp.epi_cl = kp.epi_part + (uint32_t)W * kPartRows * 128u;
  const uint32_t epi = kp.epi_cl + (uint32_t)kMaxSplit * kPartRows * 128u;
  kp.tab_bytes = (std::max(table_bytes, epi) + 127u) / 128u * 128u;
  kp.smem_ring = kp.tab_bytes;
  kp.smem_bar = kp.smem_ring + (uint32_t)W * kStages * stage;
  kp.smem_red = (kp.smem_bar + (uint32_t)W * kStages * 8u + 127u) / 128u * 128u;
  // red area: best[3][32] u64 | found[2][32] u32 | info[2][32] u32 | hyper values [12][32] f32 | gc / lp parts [W][32] f32
  //           | cluster gc partials [kMaxSplit][32] f32
  kp.smem_red_cl = kp.smem_red + 3u * 256u + 2u * 128u + 2u * 128u + 12u * 128u + (uint32_t)W * 128u;
  kp.smem_total = kp.smem_red_cl + (uint32_t)kMaxSplit * 128u;
  if (kp.smem_p2) {
    kp.smem_p2 = kp.smem_total;
    kp.smem_total += 2u * (uint32_t)T * (uint32_t)(2 + kp.ndec) * 128u;
  }
  if (kp.smem_total > kSmemMax)
    FAIL(BPLX_E_UNSUPPORTED, "problem needs %u bytes of shared memory per CTA (max %u): too many (team, confederation) pairs (%d)",
         kp.smem_total, kSmemMax, V);
#undef FAIL
  return BPLX_OK;
}

#include "plan_dynamic.inc"

}  // namespace bplx
