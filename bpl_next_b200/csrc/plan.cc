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
  } else {
    return -1;
  }
  if (K == 0) z.beta_a = z.beta_d = -1;
  *o = z;
  return d;
}

}  // namespace

int build_plan(const bplx_problem_desc& d, HostPlan* out, std::string* err) {
  const int model = d.model, M = d.num_matches, T = d.num_teams, K = d.num_covariates;
  const bool wc = model == BPLX_NEUTRAL_WC;
  const bool neutral = wc || model == BPLX_NEUTRAL;
  const int Cf = wc ? d.num_conferences : 0;
#define FAIL(code, ...)      \
  do {                       \
    *err = fmt(__VA_ARGS__); \
    return (code);           \
  } while (0)
  if (model == BPLX_DYNAMIC) FAIL(BPLX_E_UNSUPPORTED, "BPLX_DYNAMIC is not implemented by this build");
  if (model < 0 || model > BPLX_DYNAMIC) FAIL(BPLX_E_INVALID, "unknown model %d", model);
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
    if (d.weights && !(d.weights[m] >= 0.0f) )
      FAIL(BPLX_E_INVALID, "weights must be finite and non-negative (match %d)", m);
  }

  HostPlan& P = *out;
  KernelParams& kp = P.kp;
  kp.model = model;
  kp.T = T;
  kp.K = K;
  kp.Cf = Cf;
  kp.clip = model == BPLX_EXTENDED;
  kp.D = make_layout(model, T, K, Cf, &kp.off, &P.layout);

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

  // ---- number of warps -------------------------------------------------------------------------
  int W = 16;
  if (const char* e = getenv("BPLX_NWARPS")) W = atoi(e);
  if (W > kMaxWarps) W = kMaxWarps;
  if (W < 1) W = 1;
  kp.nwarps = W;

  // ---- per-match entries -----------------------------------------------------------------------
  // raw1[v][kind], raw2[v][kind]
  std::vector<std::vector<RawEntry>> raw1((size_t)V * 4), raw2((size_t)V * 4);
  std::vector<double> yexp((size_t)V * 6, 0.0);
  double w11 = 0.0, const_term = 0.0;
  for (int m = 0; m < M; m++) {
    const int h = d.home_team[m], a = d.away_team[m];
    const int hc = wc ? d.home_conf[m] : 0, ac = wc ? d.away_conf[m] : 0;
    const int hv = vmap[{h, hc}], av = vmap[{a, ac}];
    const bool nv = neutral && d.neutral_venue && d.neutral_venue[m];
    const double w = d.weights ? (double)d.weights[m] : 1.0;
    const int yh = d.home_goals[m], ya = d.away_goals[m];
    const_term -= w * (std::lgamma(yh + 1.0) + std::lgamma(ya + 1.0));
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
  P.yexp.resize(yexp.size());
  for (size_t i = 0; i < yexp.size(); i++) P.yexp[i] = (float)yexp[i];
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

  // ---- assign virtual teams to warps (longest processing time first) ---------------------------
  // All virtual teams of a team go to the same warp: the owner thread of (team, chain) is the only
  // one that read-modify-writes that team's gradient entries inside a phase.
  auto assign = [&](const std::vector<std::vector<RawEntry>>& raw, double per_entry, double per_vteam,
                    std::vector<std::vector<int>>* by_warp) {
    std::vector<double> cost(T, 0.0);
    for (int v = 0; v < V; v++) {
      size_t n = 0;
      for (int k = 0; k < 4; k++) n += raw[(size_t)v * 4 + k].size();
      if (n) cost[P.v_team[v]] += per_vteam + per_entry * (double)n;
    }
    std::vector<int> order(T);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    std::vector<double> load(W, 0.0);
    std::vector<std::vector<int>> teams(W);
    for (int t : order) {
      if (cost[t] == 0.0) continue;
      int best = 0;
      for (int w = 1; w < W; w++)
        if (load[w] < load[best]) best = w;
      load[best] += cost[t];
      teams[best].push_back(t);
    }
    by_warp->assign(W, {});
    for (int w = 0; w < W; w++) {
      std::sort(teams[w].begin(), teams[w].end());
      for (int t : teams[w])
        for (int v = P.team_vptr[t]; v < P.team_vptr[t + 1]; v++) {
          size_t n = 0;
          for (int k = 0; k < 4; k++) n += raw[(size_t)v * 4 + k].size();
          if (n) (*by_warp)[w].push_back(v);
        }
    }
  };
  std::vector<std::vector<int>> w1, w2;
  assign(raw1, 1.0, 16.0, &w1);
  assign(raw2, 3.0, 24.0, &w2);

  // ---- emit lists -------------------------------------------------------------------------------
  const uint32_t zero_row = (uint32_t)V * kRowBytes;
  P.n1 = P.n2 = 0;
  P.warp_l1.assign(W + 1, 0);
  P.warp_l2.assign(W + 1, 0);
  P.warp_e1.assign(W + 1, 0);
  P.warp_e2.assign(W + 1, 0);
  for (int w = 0; w < W; w++) {
    // phase 1
    for (int v : w1[w]) {
      size_t first = P.lists1.size();
      for (int k = 0; k < 4; k++) {
        const auto& r = raw1[(size_t)v * 4 + k];
        P.n1 += (long long)r.size();
        for (size_t lo = 0; lo < r.size(); lo += kListMax) {
          size_t hi = std::min(r.size(), lo + (size_t)kListMax);
          List L{};
          L.ent = (uint32_t)(kp.clip ? P.ent1c.size() : P.ent1.size());
          L.own_off = own_off(v, k);
          L.vteam = (uint32_t)v;
          L.kind = (uint8_t)k;
          uint32_t n = 0;
          for (size_t i = lo; i < hi; i++, n++) {
            uint32_t o = opp_base(k) + r[i].opp * kRowBytes;
            if (kp.clip)
              P.ent1c.push_back({o, (float)r[i].w, (float)r[i].wyx, (float)r[i].wyy});
            else
              P.ent1.push_back({o, (float)r[i].w});
          }
          if (!kp.clip && (n & 1)) {  // pad to an even count with a zero-row entry
            P.ent1.push_back({opp_base(k) + zero_row, 0.0f});
            n++;
          }
          L.n = n;
          P.lists1.push_back(L);
        }
      }
      if (P.lists1.size() > first) {
        P.lists1[first].flags |= kListFirst;
        P.lists1.back().flags |= kListLast;
      }
    }
    P.warp_l1[w + 1] = (int32_t)P.lists1.size();
    P.warp_e1[w + 1] = (int32_t)(kp.clip ? P.ent1c.size() : P.ent1.size());
    // phase 2
    for (int v : w2[w]) {
      size_t first = P.lists2.size();
      for (int k = 0; k < 4; k++) {
        const auto& r = raw2[(size_t)v * 4 + k];
        P.n2 += (long long)r.size();
        for (size_t lo = 0; lo < r.size(); lo += kListMax) {
          size_t hi = std::min(r.size(), lo + (size_t)kListMax);
          List L{};
          L.ent = (uint32_t)P.ent2.size();
          L.own_off = own_off(v, k);
          L.vteam = (uint32_t)v;
          L.kind = (uint8_t)k;
          uint16_t cnt[3] = {0, 0, 0};
          size_t i = lo;
          for (int c = 0; c < 3; c++) {
            for (; i < hi && r[i].cls == c; i++) {
              P.ent2.push_back({opp_base(k) + r[i].opp * kRowBytes, (float)r[i].w});
              cnt[c]++;
            }
            if (cnt[c] & 1) {
              P.ent2.push_back({opp_base(k) + zero_row, 0.0f});
              cnt[c]++;
            }
          }
          L.n_xy = cnt[0];
          L.n_x = cnt[1];
          L.n_y = cnt[2];
          L.n = (uint32_t)cnt[0] + cnt[1] + cnt[2];
          P.lists2.push_back(L);
        }
      }
      if (P.lists2.size() > first) {
        P.lists2[first].flags |= kListFirst;
        P.lists2.back().flags |= kListLast;
      }
    }
    P.warp_l2[w + 1] = (int32_t)P.lists2.size();
    P.warp_e2[w + 1] = (int32_t)P.ent2.size();
  }
  P.n1_padded = (long long)(kp.clip ? P.ent1c.size() : P.ent1.size());
  P.n2_padded = (long long)P.ent2.size();

  // ---- shared-memory carve-up --------------------------------------------------------------------
  uint32_t epi = (uint32_t)W * hyper_rows(K) * 128u;
  kp.tab_bytes = std::max(table_bytes, epi);
  kp.smem_ring = kp.tab_bytes;
  kp.smem_bar = kp.smem_ring + (uint32_t)W * kStages * kStageBytes;
  kp.smem_red = (kp.smem_bar + (uint32_t)W * kStages * 8u + 127u) / 128u * 128u;
  kp.smem_total = kp.smem_red + (uint32_t)W * kRedRows * 128u;
  if (kp.smem_total > 227u * 1024u)
    FAIL(BPLX_E_UNSUPPORTED, "problem needs %u bytes of shared memory per CTA (max %u): too many (team, confederation) pairs (%d)",
         kp.smem_total, 227u * 1024u, V);
#undef FAIL
  return BPLX_OK;
}

}  // namespace bplx
