// plan_check.cc -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
//
// Sequential, double-precision walk of the static plan that bpl_next_b200/csrc/plan.cc builds for
// the log-density kernel: tables -> entry lists -> maxima -> tau lists -> arg-max fix-up -> chain
// rule.  It exists so that the plan builder and the algebra of the factorised evaluation (which is
// NOT how the reference computes, cf. bpl/_util.py:17-93 and the `_model`s) can be checked against
// oracle/models.py on the CPU, one chain at a time.  The CUDA kernel (logdensity.cu) is the same
// walk, one lane per chain.  Never linked into the product library.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../bpl_next_b200/csrc/plan.h"

using namespace bplx;

namespace {

constexpr double kLogSqrt2Pi = 0.91893853320467274178;

struct Row { double x, y; };

const int kOwnX[4] = {eAh1, eBa1, eB0, eB0};
const int kOwnY[4] = {eBh1, eAa1, eA0, eA0};
const int kOppX[4] = {eBa1, eAh1, eA0, eA0};
const int kOppY[4] = {eAa1, eBh1, eB0, eB0};

struct Raw {  // raw slots of the walk: d/d (att, def, venue effects) per team, d/d (A - B) per virtual team
  std::vector<double> att, def, x[4], conf_v;
  double hacc = 0;  // DIXON_COLES: d/d home_advantage
};

// exponent gradients of one virtual team -> raw slots
void put_raw(const HostPlan& P, Raw& R, int v, const double g[6]) {
  const int t = P.v_team[v];
  const double gA = g[eAh1] + g[eAa1] + g[eA0], gB = g[eBh1] + g[eBa1] + g[eB0];
  R.att[t] += gA;
  R.def[t] -= gB;
  R.x[0][t] += g[eAh1];
  R.x[1][t] += g[eAa1];
  R.x[2][t] -= g[eBh1];
  R.x[3][t] -= g[eBa1];
  R.conf_v[v] += gA - gB;
  if (P.kp.ndec == 0) R.hacc += g[eAh1];
}

const ListHdr* hdr_at(const std::vector<unsigned char>& S, size_t off) {
  return reinterpret_cast<const ListHdr*>(&S[off]);
}

}  // namespace


// ---- DYNAMIC: the same walk as logdensity_dynamic.cu, one chain, double precision ----------------------------
// forward pass (gameweeks ascending: walk prefix, tables, home-list maxima) -> bounds -> backward pass (gameweeks
// descending: tables, every (gameweek, team) completed in one visit, arg-max search) -> linear fix-up corrections.
namespace {
struct DynReader {  // pieces of one warp's stream, stage by stage (pieces never straddle a stage)
  const std::vector<unsigned char>& S;
  size_t begin, end, pos;
  uint32_t stage;
  DynReader(const std::vector<unsigned char>& s, size_t b, size_t e, uint32_t st) : S(s), begin(b), end(e), pos(b), stage(st) {}
  // returns the next real header (skipping fillers); nullptr at the end of the stream
  const ListHdr* next(size_t* hoff) {
    for (;;) {
      if (pos + sizeof(ListHdr) > end) return nullptr;
      const ListHdr* L = hdr_at(S, pos);
      if (L->flags & kStageEnd) {
        pos = begin + ((pos - begin) / stage + 1) * stage;
        continue;
      }
      *hoff = pos;
      pos += sizeof(ListHdr);
      return L;
    }
  }
};

int eval_dynamic(const HostPlan& P, const double* theta, double* lp_out, double* grad, double* corr_out) {
  const KernelParams& kp = P.kp;
  const ThetaOffsets& o = kp.off;
  const int T = kp.T, K = kp.K, G = kp.G, D = kp.D, W = kp.nwarps;
  for (int i = 0; i < D; i++) grad[i] = 0.0;
  double lp = 0;
  const double mu_d = theta[o.mean_defence];
  std::vector<Row> tab(kp.tab_bytes / kRowBytes + 4, Row{0, 0});
  auto row = [&](uint32_t off) -> Row& { return tab[off / kRowBytes]; };
  std::vector<double> att((size_t)G * T, 0.0), def((size_t)G * T, 0.0);  // the kernel's workspace
  auto build = [&](int j, bool with_lp) {
    for (int t = 0; t < T; t++) {
      const int jt = j * T + t;
      if (!(P.team_flags[jt] & 1)) continue;
      double x[4];
      for (int i = 0; i < 4; i++) x[i] = theta[o.mean[i] + j] + std::exp(theta[o.log_std[i] + j]) * theta[o.dec[i] + jt];
      double ex[6];
      ex[eAh1] = att[jt] + x[0];
      ex[eBh1] = -def[jt] - x[2];
      ex[eBa1] = -def[jt] - x[3];
      ex[eAa1] = att[jt] + x[1];
      ex[eA0] = att[jt];
      ex[eB0] = -def[jt];
      if (kp.has1) {
        row(kp.tabP1 + t * kRowBytes) = Row{std::exp(ex[eAh1]), std::exp(ex[eBh1])};
        row(kp.tabQ1 + t * kRowBytes) = Row{std::exp(ex[eBa1]), std::exp(ex[eAa1])};
      }
      if (kp.has0) row(kp.tabP0 + t * kRowBytes) = Row{std::exp(ex[eA0]), std::exp(ex[eB0])};
      if (with_lp)
        for (int e = 0; e < 6; e++) lp += P.yexp[(size_t)jt * 6 + e] * ex[e];
    }
  };
  // ---- forward pass ------------------------------------------------------------------------------------------
  {
    std::vector<double> a(T), d(T);
    for (int t = 0; t < T; t++) {
      a[t] = 0;
      d[t] = mu_d;
      for (int k = 0; k < K; k++) {
        a[t] += P.Xs[(size_t)t * K + k] * theta[o.beta_a + k];
        d[t] += P.Xs[(size_t)t * K + k] * theta[o.beta_d + k];
      }
    }
    for (int j = 0; j < G; j++)
      for (int t = 0; t < T; t++) {
        const int jt = j * T + t;
        if (!kp.as_written) {
          a[t] += theta[o.za + jt] * std::exp(theta[o.log_std_attack + j]);
          d[t] += theta[o.zd + jt] * std::exp(theta[o.log_std_defence + j]);
          att[jt] = a[t];
          def[jt] = d[t];
        }
      }
  }
  double best[3] = {0, 0, 0};
  size_t best_hdr[3] = {0, 0, 0};
  {
    std::vector<DynReader> rd;
    for (int w = 0; w < W; w++) rd.emplace_back(P.stream1, P.warp_b1[w], P.warp_b1[w + 1], kp.stage_bytes);
    for (int j = 0; j < G; j++) {
      build(j, false);
      for (int w = 0; w < W; w++) {
        size_t hoff;
        const ListHdr* Mk = rd[w].next(&hoff);
        if (!Mk || !(Mk->flags & kGwFirst) || Mk->vteam != j) return -130;
        for (int p = 0; p < Mk->n0; p++) {
          const ListHdr* L = rd[w].next(&hoff);
          if (!L || (L->flags & (kGwFirst | kPhase2)) || L->vteam != j || (L->n0 & 1)) return -131;
          if (!(L->kind == kH1 || L->kind == kH0) || (int)(L->team % W) != w) return -132;
          if ((hoff - rd[w].begin) / kp.stage_bytes != (rd[w].pos + L->n0 * sizeof(Entry) - 1 - rd[w].begin) / kp.stage_bytes)
            return -135;  // a piece straddles a stage
          Row own = row(L->own_off);
          if (L->kind >= kH0) std::swap(own.x, own.y);
          double m1 = 0, m2 = 0, m3 = 0;
          for (uint32_t i = 0; i < L->n0; i++, rd[w].pos += sizeof(Entry)) {
            const Entry& e = *reinterpret_cast<const Entry*>(&P.stream1[rd[w].pos]);
            const Row& op = row(e.off);
            m1 = std::max(m1, op.x);
            m2 = std::max(m2, op.y);
            m3 = std::max(m3, op.x * op.y);
          }
          const double v[3] = {own.x * m1, own.y * m2, own.x * own.y * m3};
          for (int q = 0; q < 3; q++)
            if (v[q] > best[q] || (v[q] == best[q] && v[q] > 0 && hoff > best_hdr[q])) best[q] = v[q], best_hdr[q] = hoff;
        }
      }
    }
    for (int w = 0; w < W; w++) {
      size_t hoff;
      if (rd[w].next(&hoff) != nullptr) return -133;
    }
  }
  const double Lam = std::max(best[0], best[1]);
  const double LB = -1.0 / Lam, UB = std::min(1.0 / best[2], 1.0);
  const double r = 1.0 / (1.0 + std::exp(-theta[o.raw]));
  const double cc = LB + r * (UB - LB);
  *corr_out = cc;
  const int qlam = best[0] >= best[1] ? 0 : 1;
  struct Hit { int j, own, opp, kind; bool ok; };
  Hit hit[2] = {{0, 0, 0, 0, false}, {0, 0, 0, 0, false}};
  const int want_gw[2] = {(int)hdr_at(P.stream1, best_hdr[qlam])->vteam,
                          best[2] > 1.0 ? (int)hdr_at(P.stream1, best_hdr[2])->vteam : -1};
  auto find = [&](int which) {  // with the tables of the arg-max piece's gameweek resident
    const int q = which == 0 ? qlam : 2;
    const ListHdr& L = *hdr_at(P.stream1, best_hdr[q]);
    Row own = row(L.own_off);
    if (L.kind >= kH0) std::swap(own.x, own.y);
    for (uint32_t i = 0; i < L.n0; i++) {
      const uint32_t off = *reinterpret_cast<const uint32_t*>(&P.stream1[best_hdr[q] + 16 + i * sizeof(Entry)]);
      const Row& op = row(off);
      const double val = q == 0 ? own.x * op.x : q == 1 ? own.y * op.y : own.x * own.y * (op.x * op.y);
      if (val == best[q]) {
        hit[which] = Hit{L.vteam, L.team, (int)((off - (L.kind == kH1 ? kp.tabQ1 : kp.tabP0)) / kRowBytes), L.kind, true};
        break;
      }
    }
  };
  // ---- backward pass ---------------------------------------------------------------------------------------------
  double Gc = 0;
  std::vector<double> s_att(T, 0.0), s_def(T, 0.0);
  std::vector<int> visited((size_t)G * T, 0);
  {
    std::vector<DynReader> rd;
    for (int w = 0; w < W; w++) rd.emplace_back(P.stream2, P.warp_b2[w], P.warp_b2[w + 1], kp.stage_bytes);
    for (int j = G - 1; j >= 0; j--) {
      build(j, true);
      for (int which = 0; which < 2; which++)
        if (want_gw[which] == j) find(which);
      const double sig_a = std::exp(theta[o.log_std_attack + j]), sig_d = std::exp(theta[o.log_std_defence + j]);
      double a_ls_a = 0, a_ls_d = 0, a_mu[4] = {0, 0, 0, 0}, a_ls[4] = {0, 0, 0, 0}, sig[4];
      for (int i = 0; i < 4; i++) sig[i] = std::exp(theta[o.log_std[i] + j]);
      for (int w = 0; w < W; w++) {
        size_t hoff;
        const ListHdr* Mk = rd[w].next(&hoff);
        if (!Mk || !(Mk->flags & kGwFirst) || Mk->vteam != j) return -120;
        double g[6] = {0, 0, 0, 0, 0, 0};
        int cur_team = -1;
        for (int p = 0; p < Mk->n0; p++) {
          const ListHdr* L = rd[w].next(&hoff);
          if (!L || (L->flags & kGwFirst) || L->vteam != j || ((L->n0 | L->n1 | L->n2) & 1)) return -121;
          if ((int)(L->team % W) != w) return -122;
          if (L->flags & kTeamFirst) {
            if (cur_team != -1) return -123;
            cur_team = L->team;
            for (int e = 0; e < 6; e++) g[e] = 0;
          }
          if (cur_team != (int)L->team) return -124;
          const size_t nent = (size_t)L->n0 + ((L->flags & kPhase2) ? L->n1 + L->n2 : 0);
          if (!(L->flags & kPhase2) && (L->n1 | L->n2)) return -125;
          if (nent && (hoff - rd[w].begin) / kp.stage_bytes != (rd[w].pos + nent * sizeof(Entry) - 1 - rd[w].begin) / kp.stage_bytes)
            return -126;
          Row own = row(L->own_off);
          if (L->kind >= kH0) std::swap(own.x, own.y);
          const bool home = L->kind == kH1 || L->kind == kH0;
          double gl[6] = {0, 0, 0, 0, 0, 0};
          if (!(L->flags & kPhase2)) {
            double ax = 0, ay = 0;
            for (uint32_t i = 0; i < L->n0; i++, rd[w].pos += sizeof(Entry)) {
              const Entry& e = *reinterpret_cast<const Entry*>(&P.stream2[rd[w].pos]);
              const Row& op = row(e.off);
              ax += e.w * op.x;
              ay += e.w * op.y;
            }
            const double SX = own.x * ax, SY = own.y * ay;
            lp -= 0.5 * (SX + SY);
            gl[kOwnX[L->kind]] -= SX;
            gl[kOwnY[L->kind]] -= SY;
          } else {
            double gx = 0, gy = 0, lt = 0, dG = 0;
            const uint32_t n[3] = {L->n0, L->n1, L->n2};
            for (int cls = 0; cls < 3; cls++)
              for (uint32_t k = 0; k < n[cls]; k++, rd[w].pos += sizeof(Entry)) {
                const Entry& e = *reinterpret_cast<const Entry*>(&P.stream2[rd[w].pos]);
                if (e.w == 0.0f) continue;
                if (cls == 0) {
                  const Row& op = row(e.off);
                  const double X = own.x * op.x, Y = own.y * op.y;
                  const double tau = 1.0 - cc * X * Y, q = e.w * X * Y / tau;
                  lt += e.w * std::log(tau);
                  dG -= q;
                  gx -= cc * q;
                  gy -= cc * q;
                } else if (cls == 1) {
                  const double X = own.x * row(e.off).x;
                  const double tau = 1.0 + cc * X, q = e.w * X / tau;
                  lt += e.w * std::log(tau);
                  dG += q;
                  gx += cc * q;
                } else {
                  if ((e.off & 7u) != 4u) return -127;  // `off` addresses the .y float
                  const double Y = own.y * row(e.off).y;
                  const double tau = 1.0 + cc * Y, q = e.w * Y / tau;
                  lt += e.w * std::log(tau);
                  dG += q;
                  gy += cc * q;
                }
              }
            if (home) {
              lp += lt;
              Gc += dG;
            }
            gl[kOwnX[L->kind]] += gx;
            gl[kOwnY[L->kind]] += gy;
          }
          for (int e = 0; e < 6; e++) g[e] += gl[e];
          if (L->flags & kTeamLast) {  // complete (gameweek j, team t)
            const int t = L->team, jt = j * T + t;
            if (visited[jt]++) return -128;
            const float* ys = &P.yteam[(size_t)jt * 8];
            const double ra = g[eAh1] + g[eAa1] + g[eA0] + ys[0], rdd = -(g[eBh1] + g[eBa1] + g[eB0]) + ys[1];
            const double rx[4] = {g[eAh1] + ys[2], g[eAa1] + ys[3], -g[eBh1] + ys[4], -g[eBa1] + ys[5]};
            s_att[t] += ra;
            s_def[t] += rdd;
            const double sa = kp.as_written ? 0.0 : s_att[t], sd = kp.as_written ? 0.0 : s_def[t];
            const double za = theta[o.za + jt], zd = theta[o.zd + jt];
            const double u = 1.0 / (1.0 + std::exp(-theta[o.u + jt]));
            const double rho = 2.0 * u - 1.0, s2 = 1.0 - rho * rho;
            const double e = zd - rho * za;
            lp += -0.5 * za * za - 0.5 * e * e / s2 - 0.5 * std::log(s2) + 2.0 * std::log(u) + 4.0 * std::log(1.0 - u);
            const double a_rho = e * za / s2 - rho * e * e / (s2 * s2) + rho / s2;
            grad[o.u + jt] = 2.0 - 6.0 * u + a_rho * 2.0 * u * (1.0 - u);
            grad[o.za + jt] = -za + rho * e / s2 + sig_a * sa;
            grad[o.zd + jt] = -e / s2 + sig_d * sd;
            a_ls_a += sig_a * za * sa;
            a_ls_d += sig_d * zd * sd;
            for (int i = 0; i < 4; i++) {
              const double dec = theta[o.dec[i] + jt];
              lp -= 0.5 * dec * dec;
              grad[o.dec[i] + jt] = -dec + sig[i] * rx[i];
              a_mu[i] += rx[i];
              a_ls[i] += sig[i] * dec * rx[i];
            }
            cur_team = -1;
          }
        }
        if (cur_team != -1) return -129;
      }
      lp += -0.5 * sig_a * sig_a + theta[o.log_std_attack + j] - 0.5 * sig_d * sig_d + theta[o.log_std_defence + j];
      grad[o.log_std_attack + j] = -sig_a * sig_a + 1.0 + a_ls_a;
      grad[o.log_std_defence + j] = -sig_d * sig_d + 1.0 + a_ls_d;
      for (int i = 0; i < 4; i++) {
        const double z = (theta[o.mean[i] + j] - ((i & 1) ? -0.1 : 0.1)) * 5.0;
        lp += -0.5 * z * z - 0.5 * sig[i] * sig[i] + theta[o.log_std[i] + j];
        grad[o.mean[i] + j] = -z * 5.0 + a_mu[i];
        grad[o.log_std[i] + j] = -sig[i] * sig[i] + 1.0 + a_ls[i];
      }
    }
    for (int w = 0; w < W; w++) {
      size_t hoff;
      if (rd[w].next(&hoff) != nullptr) return -136;
    }
    for (size_t i = 0; i < visited.size(); i++)
      if (visited[i] != 1) return -137;
  }
  if (!hit[0].ok || (best[2] > 1.0 && !hit[1].ok)) return -100;
  lp += kp.w11 * std::log(1.0 - cc);
  Gc -= kp.w11 / (1.0 - cc);
  // mean_defence, covariate coefficients: the whole walk's sums
  double g_md = -mu_d;
  lp -= 0.5 * mu_d * mu_d;
  std::vector<double> g_ba(K, 0.0), g_bd(K, 0.0);
  if (!kp.as_written)
    for (int t = 0; t < T; t++) {
      g_md += s_def[t];
      for (int k = 0; k < K; k++) {
        g_ba[k] += P.Xs[(size_t)t * K + k] * s_att[t];
        g_bd[k] += P.Xs[(size_t)t * K + k] * s_def[t];
      }
    }
  for (int k = 0; k < K; k++) {
    const double ba = theta[o.beta_a + k], bd = theta[o.beta_d + k];
    lp -= 0.5 * (ba * ba + bd * bd);
    g_ba[k] -= ba;
    g_bd[k] -= bd;
  }
  // ---- fix-up: d corr_coef / d eta of the arg-max matches, applied as a linear correction -----
  for (int which = 0; which < 2; which++) {
    if (!hit[which].ok) continue;
    double vx, vy;
    if (which == 0) {
      const double wgt = Gc * (1.0 - r) / Lam;
      vx = qlam == 0 ? wgt : 0.0;
      vy = qlam == 1 ? wgt : 0.0;
    } else {
      vx = vy = -Gc * r / best[2];
    }
    const Hit& H = hit[which];
    for (int side = 0; side < 2; side++) {
      double gl[6] = {0, 0, 0, 0, 0, 0};
      gl[side == 0 ? kOwnX[H.kind] : kOppX[H.kind]] += vx;
      gl[side == 0 ? kOwnY[H.kind] : kOppY[H.kind]] += vy;
      const int t = side == 0 ? H.own : H.opp, js = H.j;
      const double dra = gl[eAh1] + gl[eAa1] + gl[eA0], drd = -(gl[eBh1] + gl[eBa1] + gl[eB0]);
      const double drx[4] = {gl[eAh1], gl[eAa1], -gl[eBh1], -gl[eBa1]};
      for (int i = 0; i < 4; i++) {
        const double sg = std::exp(theta[o.log_std[i] + js]);
        grad[o.dec[i] + js * T + t] += sg * drx[i];
        grad[o.mean[i] + js] += drx[i];
        grad[o.log_std[i] + js] += sg * theta[o.dec[i] + js * T + t] * drx[i];
      }
      if (!kp.as_written) {
        for (int j = 0; j <= js; j++) {
          const double sa = std::exp(theta[o.log_std_attack + j]), sd = std::exp(theta[o.log_std_defence + j]);
          grad[o.za + j * T + t] += sa * dra;
          grad[o.zd + j * T + t] += sd * drd;
          grad[o.log_std_attack + j] += sa * theta[o.za + j * T + t] * dra;
          grad[o.log_std_defence + j] += sd * theta[o.zd + j * T + t] * drd;
        }
        g_md += drd;
        for (int k = 0; k < K; k++) {
          g_ba[k] += P.Xs[(size_t)t * K + k] * dra;
          g_bd[k] += P.Xs[(size_t)t * K + k] * drd;
        }
      }
    }
  }
  grad[o.mean_defence] = g_md;
  for (int k = 0; k < K; k++) {
    grad[o.beta_a + k] = g_ba[k];
    grad[o.beta_d + k] = g_bd[k];
  }
  lp += std::log(r) + std::log(1.0 - r);  // Uniform(0,1): Jacobian only
  grad[o.raw] = (1.0 - 2.0 * r) + Gc * r * (1.0 - r) * (UB - LB);
  lp += kp.const_term;
  *lp_out = lp;
  return 0;
}
}  // namespace

extern "C" int bplx_plancheck_eval2(const bplx_problem_desc* desc, int split_idx, const double* theta, double* lp_out,
                                    double* grad, double* corr_out, char* errbuf, int errlen);
extern "C" int bplx_plancheck_eval(const bplx_problem_desc* desc, const double* theta, double* lp_out, double* grad,
                                   double* corr_out, char* errbuf, int errlen) {
  return bplx_plancheck_eval2(desc, 0, theta, lp_out, grad, corr_out, errbuf, errlen);
}

// split_idx i walks the streams built for 2^i CTAs per chain group (nwarps << i virtual warps)
extern "C" int bplx_plancheck_eval2(const bplx_problem_desc* desc, int split_idx, const double* theta, double* lp_out,
                                    double* grad, double* corr_out, char* errbuf, int errlen) {
  HostPlan P;
  std::string err;
  int rc = build_plan(*desc, &P, &err);
  if (rc != BPLX_OK) {
    if (errbuf && errlen > 0) snprintf(errbuf, errlen, "%s", err.c_str());
    return rc;
  }
  if (split_idx < 0 || split_idx >= kNumSplits) return -140;
  if (split_idx > 0) {  // walk the split's streams in place of split 1
    if (P.kp.model == BPLX_DYNAMIC) return -141;
    P.stream1 = P.more[split_idx - 1].stream1;
    P.stream2 = P.more[split_idx - 1].stream2;
    P.warp_b1 = P.more[split_idx - 1].warp_b1;
    P.warp_b2 = P.more[split_idx - 1].warp_b2;
    P.kp.nwarps <<= split_idx;
  }
  const KernelParams& kp = P.kp;
  if (kp.model == BPLX_DYNAMIC) return eval_dynamic(P, theta, lp_out, grad, corr_out);
  const ThetaOffsets& o = kp.off;
  const int T = kp.T, K = kp.K, V = kp.V, Cf = kp.Cf, D = kp.D, ndec = kp.ndec;
  for (int i = 0; i < D; i++) grad[i] = 0.0;
  const bool dc = kp.model == BPLX_DIXON_COLES;
  const bool has_rho = !dc;
  const size_t esz1 = kp.clip ? sizeof(EntryClip) : sizeof(Entry);
  double lp = 0;
  // ---- hypers -----------------------------------------------------------------------------------
  const double mu_d = theta[o.mean_defence];
  const double sig_a = std::exp(theta[o.log_std_attack]), sig_d = std::exp(theta[o.log_std_defence]);
  double mu[4], sig[4];
  for (int i = 0; i < 4; i++) {
    mu[i] = o.mean[i] >= 0 ? theta[o.mean[i]] : 0.0;
    sig[i] = o.log_std[i] >= 0 ? std::exp(theta[o.log_std[i]]) : 0.0;
  }
  const double u = has_rho ? 1.0 / (1.0 + std::exp(-theta[o.u])) : 0.5;
  const double rho = has_rho ? 2.0 * u - 1.0 : 0.0, s2 = 1.0 - rho * rho;
  // ---- prologue: tables, static y part ------------------------------------------------------------
  std::vector<Row> tab(kp.tab_bytes / kRowBytes + 4, Row{0, 0});  // one Row per table row
  auto row = [&](uint32_t off) -> Row& { return tab[off / kRowBytes]; };
  for (int t = 0; t < T; t++) {
    double am = 0, dm = mu_d;
    for (int k = 0; k < K; k++) {
      am += P.Xs[(size_t)t * K + k] * theta[o.beta_a + k];
      dm += P.Xs[(size_t)t * K + k] * theta[o.beta_d + k];
    }
    const double att = am + theta[o.za + t] * sig_a, def = dm + theta[o.zd + t] * sig_d;
    double x[4] = {dc ? mu[0] : 0.0, 0, 0, 0};
    for (int i = 0; i < ndec; i++) x[i] = mu[i] + sig[i] * theta[o.dec[i] + t];
    for (int v = P.team_vptr[t]; v < P.team_vptr[t + 1]; v++) {
      const double cf = Cf > 0 ? theta[o.conf + P.v_conf[v]] : 0.0;
      double ex[6];
      ex[eAh1] = att + x[0] + cf;
      ex[eBh1] = -def - x[2] - cf;
      ex[eBa1] = -def - x[3] - cf;
      ex[eAa1] = att + x[1] + cf;
      ex[eA0] = att + cf;
      ex[eB0] = -def - cf;
      if (kp.has1) {
        row(kp.tabP1 + v * kRowBytes) = Row{std::exp(ex[eAh1]), std::exp(ex[eBh1])};
        row(kp.tabQ1 + v * kRowBytes) = Row{std::exp(ex[eBa1]), std::exp(ex[eAa1])};
      }
      if (kp.has0) row(kp.tabP0 + v * kRowBytes) = Row{std::exp(ex[eA0]), std::exp(ex[eB0])};
      for (int e = 0; e < 6; e++) lp += P.yexp[(size_t)v * 6 + e] * ex[e];
    }
  }
  Raw R;
  R.att.assign(T, 0.0);
  R.def.assign(T, 0.0);
  for (auto& v : R.x) v.assign(T, 0.0);
  R.conf_v.assign(V, 0.0);
  // ---- phase 1: walk every warp's stream ------------------------------------------------------------
  double best[3] = {0, 0, 0};  // max X, max Y, max XY over home-role lists
  size_t best_hdr[3] = {0, 0, 0};
  std::vector<int> seen_first(T, 0), seen_last(T, 0);
  for (int w = 0; w < kp.nwarps; w++) {
    size_t pos = P.warp_b1[w];
    const size_t end = P.warp_b1[w + 1], wstart = pos;
    double g[6] = {0, 0, 0, 0, 0, 0};
    while (pos < end) {
      {  // dead stage tail
        const size_t in_stage = (pos - wstart) % kp.stage_bytes;
        if (kp.stage_bytes - in_stage < kp.min_piece1) { pos += kp.stage_bytes - in_stage; continue; }
      }
      const size_t hoff = pos;
      const ListHdr& L = *hdr_at(P.stream1, pos);
      pos += sizeof(ListHdr);
      if (L.flags & kTeamFirst) {
        memset(g, 0, sizeof g);
        if (seen_first[L.team]++) return -110;  // a team must be owned by exactly one warp
      }
      if (P.v_team[L.vteam] != L.team) return -111;
      Row own = row(L.own_off);
      if (L.kind >= kH0) std::swap(own.x, own.y);
      const bool home = L.kind == kH1 || L.kind == kH0;
      double gx = 0, gy = 0;
      if (!kp.clip) {
        if (L.n0 & 1) return -112;
        double ax = 0, ay = 0, m1 = 0, m2 = 0, m3 = 0;
        for (uint32_t i = 0; i < L.n0; i++, pos += esz1) {
          const Entry& e = *reinterpret_cast<const Entry*>(&P.stream1[pos]);
          const Row& op = row(e.off);
          ax += e.w * op.x;
          ay += e.w * op.y;
          m1 = std::max(m1, op.x);
          m2 = std::max(m2, op.y);
          m3 = std::max(m3, op.x * op.y);
        }
        const double SX = own.x * ax, SY = own.y * ay;
        lp -= 0.5 * (SX + SY);
        gx = -SX;
        gy = -SY;
        if (home) {
          const double v[3] = {own.x * m1, own.y * m2, own.x * own.y * m3};
          for (int q = 0; q < 3; q++)
            if (v[q] > best[q]) best[q] = v[q], best_hdr[q] = hoff;
        }
      } else {
        for (uint32_t i = 0; i < L.n0; i++, pos += esz1) {
          const EntryClip& e = *reinterpret_cast<const EntryClip*>(&P.stream1[pos]);
          const Row& op = row(e.off);
          const double X = own.x * op.x, Y = own.y * op.y;
          const double Xc = std::min(X, 15.0), Yc = std::min(Y, 15.0);
          if (home) {
            lp += e.wyx * std::log(Xc) - e.w * Xc + e.wyy * std::log(Yc) - e.w * Yc;
            const double v[3] = {Xc, Yc, Xc * Yc};
            for (int q = 0; q < 3; q++)
              if (v[q] > best[q]) best[q] = v[q], best_hdr[q] = hoff;
          }
          if (X < 15.0) gx += e.wyx - e.w * X;
          if (Y < 15.0) gy += e.wyy - e.w * Y;
        }
      }
      double gl[6] = {0, 0, 0, 0, 0, 0};
      gl[kOwnX[L.kind]] += gx;
      gl[kOwnY[L.kind]] += gy;
      put_raw(P, R, L.vteam, gl);
      if (L.flags & kTeamLast) seen_last[L.team]++;
    }
    if (pos != end) return -113;
  }
  for (int t = 0; t < T; t++)
    if (seen_first[t] != seen_last[t] || seen_first[t] != (P.team_flags[t] & 1)) return -114;
  // ---- bounds and corr_coef (bpl/_util.py:17-31) ---------------------------------------------------
  const double Lam = std::max(best[0], best[1]);
  const double LB = -1.0 / Lam;
  const double UB = std::min(1.0 / best[2], 1.0);
  const double r = 1.0 / (1.0 + std::exp(-theta[o.raw]));
  const double cc = LB + r * (UB - LB);
  *corr_out = cc;
  // ---- phase 2: tau terms ----------------------------------------------------------------------------
  double Gc = 0;
  for (int w = 0; w < kp.nwarps; w++) {
    size_t pos = P.warp_b2[w];
    const size_t end = P.warp_b2[w + 1], wstart = pos;
    while (pos < end) {
      {
        const size_t in_stage = (pos - wstart) % kp.stage_bytes;
        if (kp.stage_bytes - in_stage < kp.min_piece2) { pos += kp.stage_bytes - in_stage; continue; }
      }
      const ListHdr& L = *hdr_at(P.stream2, pos);
      pos += sizeof(ListHdr);
      if ((L.n0 | L.n1 | L.n2) & 1) return -120;
      Row own = row(L.own_off);
      if (L.kind >= kH0) std::swap(own.x, own.y);
      const bool home = L.kind == kH1 || L.kind == kH0;
      double gx = 0, gy = 0, lt = 0, dG = 0;
      const uint32_t n[3] = {L.n0, L.n1, L.n2};
      for (int cls = 0; cls < 3; cls++) {
        for (uint32_t k = 0; k < n[cls]; k++, pos += sizeof(Entry)) {
          const Entry& e = *reinterpret_cast<const Entry*>(&P.stream2[pos]);
          if (e.w == 0.0f) continue;  // padding
          const Row& op = row(e.off);
          const double Xr = own.x * op.x, Yr = own.y * op.y;
          const double X = kp.clip ? std::min(Xr, 15.0) : Xr, Y = kp.clip ? std::min(Yr, 15.0) : Yr;
          const bool xfree = !kp.clip || Xr < 15.0, yfree = !kp.clip || Yr < 15.0;  // a clipped rate has zero derivative
          if (cls == 0) {
            const double tau = 1.0 - cc * X * Y, q = e.w * X * Y / tau;
            lt += e.w * std::log(tau);
            dG -= q;
            if (xfree) gx -= cc * q;
            if (yfree) gy -= cc * q;
          } else if (cls == 1) {
            const double tau = 1.0 + cc * X, q = e.w * X / tau;
            lt += e.w * std::log(tau);
            dG += q;
            if (xfree) gx += cc * q;
          } else {
            const double tau = 1.0 + cc * Y, q = e.w * Y / tau;
            lt += e.w * std::log(tau);
            dG += q;
            if (yfree) gy += cc * q;
          }
        }
      }
      if (home) {
        lp += lt;
        Gc += dG;
      }
      double gl[6] = {0, 0, 0, 0, 0, 0};
      gl[kOwnX[L.kind]] += gx;
      gl[kOwnY[L.kind]] += gy;
      put_raw(P, R, L.vteam, gl);
    }
    if (pos != end) return -121;
  }
  lp += kp.w11 * std::log(1.0 - cc);
  Gc -= kp.w11 / (1.0 - cc);
  // ---- arg-max fix-up -----------------------------------------------------------
  // dc/dLB = 1 - r, dLB/dLam = 1/Lam^2, dLam/d eta = Lam  (zero when the max is a clipped rate)
  auto find = [&](int q, int* own_v, int* opp_v, int* kind, double* Xr, double* Yr) {
    const ListHdr& L = *hdr_at(P.stream1, best_hdr[q]);
    Row own = row(L.own_off);
    if (L.kind >= kH0) std::swap(own.x, own.y);
    *own_v = (int)L.vteam;
    *kind = L.kind;
    *opp_v = -1;
    for (uint32_t i = 0; i < L.n0; i++) {
      const uint32_t off = *reinterpret_cast<const uint32_t*>(&P.stream1[best_hdr[q] + 16 + i * esz1]);
      const Row& op = row(off);
      double X = own.x * op.x, Y = own.y * op.y;
      double Xc = kp.clip ? std::min(X, 15.0) : X, Yc = kp.clip ? std::min(Y, 15.0) : Y;
      double val = q == 0 ? Xc : q == 1 ? Yc : (kp.clip ? Xc * Yc : own.x * own.y * (op.x * op.y));
      if (val == best[q]) {
        const uint32_t base = L.kind == kH1 ? kp.tabQ1 : kp.tabP0;
        *opp_v = (int)((off - base) / kRowBytes);
        *Xr = X;
        *Yr = Y;
        return;
      }
    }
  };
  {
    const int q = best[0] >= best[1] ? 0 : 1;
    int ov, pv, kind;
    double Xr, Yr;
    find(q, &ov, &pv, &kind, &Xr, &Yr);
    if (pv < 0) return -100;
    const double rate = q == 0 ? Xr : Yr;
    if (!kp.clip || rate < 15.0) {
      const double wgt = Gc * (1.0 - r) / Lam;
      double g1[6] = {0}, g2[6] = {0};
      g1[q == 0 ? kOwnX[kind] : kOwnY[kind]] = wgt;
      g2[q == 0 ? kOppX[kind] : kOppY[kind]] = wgt;
      put_raw(P, R, ov, g1);
      put_raw(P, R, pv, g2);
    }
  }
  if (best[2] > 1.0) {
    int ov, pv, kind;
    double Xr, Yr;
    find(2, &ov, &pv, &kind, &Xr, &Yr);
    if (pv < 0) return -101;
    const double wgt = -Gc * r / best[2];
    double g1[6] = {0}, g2[6] = {0};
    if (!kp.clip || Xr < 15.0) g1[kOwnX[kind]] += wgt, g2[kOppX[kind]] += wgt;
    if (!kp.clip || Yr < 15.0) g1[kOwnY[kind]] += wgt, g2[kOppY[kind]] += wgt;
    put_raw(P, R, ov, g1);
    put_raw(P, R, pv, g2);
  }
  // ---- team pass: raw slots + static parts -> parameter gradients, priors, hyper sums ------------------
  double acc[kAccRows] = {0};
  std::vector<double> a_ba(K, 0.0), a_bd(K, 0.0);
  for (int t = 0; t < T; t++) {
    const double za = theta[o.za + t], zd = theta[o.zd + t];
    const float* ys = &P.yteam[(size_t)t * 8];
    const double ra = R.att[t] + ys[0], rd = R.def[t] + ys[1];
    if (has_rho) {
      const double e = zd - rho * za;
      lp += -0.5 * za * za - 0.5 * e * e / s2;
      grad[o.za + t] = -za + rho * e / s2 + sig_a * ra;
      grad[o.zd + t] = -e / s2 + sig_d * rd;
      acc[12] += e * za / s2 - rho * e * e / (s2 * s2) + rho / s2;
    } else {
      lp += -0.5 * za * za - 0.5 * zd * zd;
      grad[o.za + t] = -za + sig_a * ra;
      grad[o.zd + t] = -zd + sig_d * rd;
    }
    acc[2] += sig_a * za * ra;
    acc[3] += sig_d * zd * rd;
    acc[1] += rd;
    if (dc) acc[4] += ys[2];
    for (int i = 0; i < ndec; i++) {
      const double dec = theta[o.dec[i] + t], rx = R.x[i][t] + ys[2 + i];
      lp -= 0.5 * dec * dec;
      grad[o.dec[i] + t] = -dec + sig[i] * rx;
      acc[4 + i] += rx;
      acc[8 + i] += sig[i] * dec * rx;
    }
    for (int k = 0; k < K; k++) {
      a_ba[k] += P.Xs[(size_t)t * K + k] * ra;
      a_bd[k] += P.Xs[(size_t)t * K + k] * rd;
    }
  }
  if (dc) acc[4] += R.hacc;
  for (int k = 0; k < Cf; k++) {
    const double cf = theta[o.conf + k];
    double s = P.yconf[k] - cf;
    lp -= 0.5 * cf * cf;
    for (int j = P.conf_vptr[k]; j < P.conf_vptr[k + 1]; j++) s += R.conf_v[P.conf_vlist[j]];
    grad[o.conf + k] = s;
  }
  for (int k = 0; k < K; k++) {
    const double ba = theta[o.beta_a + k], bd = theta[o.beta_d + k];
    lp -= 0.5 * (ba * ba + bd * bd);
    grad[o.beta_a + k] = -ba + a_ba[k];
    grad[o.beta_d + k] = -bd + a_bd[k];
  }
  // ---- hyper priors, Jacobians (normalising constants are in const_term) ---------------------------------
  for (int h = 0; h < kp.nhyper; h++) {
    const HyperDesc& hd = kp.hyper[h];
    const double x = theta[hd.off];
    if (hd.kind == 0) {
      const double z = (x - hd.loc) * hd.inv_scale;
      lp -= 0.5 * z * z;
      grad[hd.off] = -z * hd.inv_scale + acc[hd.row];
    } else {
      const double z = std::exp(x) * hd.inv_scale;
      lp += -0.5 * z * z + x;
      grad[hd.off] = -z * z + 1.0 + acc[hd.row];
    }
  }
  if (has_rho) {  // u ~ Beta(2, 4) + sigmoid Jacobian; rho = 2u - 1
    lp += -0.5 * T * std::log(s2) + 2.0 * std::log(u) + 4.0 * std::log(1.0 - u);
    grad[o.u] = 2.0 - 6.0 * u + acc[12] * 2.0 * u * (1.0 - u);
  }
  // corr_coef_raw ~ Beta(2, 2) + Jacobian; c = LB + r (UB - LB)
  lp += 2.0 * (std::log(r) + std::log(1.0 - r));
  grad[o.raw] = 2.0 * (1.0 - 2.0 * r) + Gc * r * (1.0 - r) * (UB - LB);
  lp += kp.const_term;
  (void)kLogSqrt2Pi;
  *lp_out = lp;
  return 0;
}

// the flat layout string and D the product reports for a problem (no GPU needed)
extern "C" int bplx_plancheck_layout(const bplx_problem_desc* desc, char* buf, int buflen) {
  HostPlan P;
  std::string err;
  int rc = build_plan(*desc, &P, &err);
  if (rc != BPLX_OK) return rc;
  snprintf(buf, buflen, "%s", P.layout.c_str());
  return P.kp.D;
}
