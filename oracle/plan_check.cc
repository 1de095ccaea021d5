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
constexpr double kLog2 = 0.69314718055994530942;

struct Row { double x, y; };

struct Ctx {
  const HostPlan* P;
  const KernelParams* kp;
  const double* th;
  double* grad;
  // constrained hypers
  double mu_d, sig_a, sig_d, mu[4], sig[4], rho, s2;
  double beta_a[kMaxCov], beta_d[kMaxCov];
  // hyper gradient accumulators
  double a_mu_d = 0, a_ls_a = 0, a_ls_d = 0, a_mu[4] = {0, 0, 0, 0}, a_ls[4] = {0, 0, 0, 0}, a_rho = 0;
  double a_ba[kMaxCov] = {0}, a_bd[kMaxCov] = {0};
  std::vector<double> a_conf;
  double lp = 0;
};

const int kOwnX[4] = {eAh1, eBa1, eB0, eB0};
const int kOwnY[4] = {eBh1, eAa1, eA0, eA0};
const int kOppX[4] = {eBa1, eAh1, eA0, eA0};
const int kOppY[4] = {eAa1, eBh1, eB0, eB0};

// exponent gradients of virtual team v -> raw parameter gradients
void apply_vteam(Ctx& c, int v, const double g[6]) {
  const KernelParams& kp = *c.kp;
  const ThetaOffsets& o = kp.off;
  const int t = c.P->v_team[v];
  const double gA = g[eAh1] + g[eAa1] + g[eA0];
  const double gB = g[eBh1] + g[eBa1] + g[eB0];
  const double g_att = gA, g_def = -gB;
  if (kp.model == BPLX_NEUTRAL_WC) c.a_conf[c.P->v_conf[v]] += gA - gB;
  const double za = c.th[o.za + t], zd = c.th[o.zd + t];
  c.grad[o.za + t] += c.sig_a * g_att;
  c.a_ls_a += c.sig_a * za * g_att;
  c.grad[o.zd + t] += c.sig_d * g_def;
  c.a_ls_d += c.sig_d * zd * g_def;
  c.a_mu_d += g_def;
  for (int k = 0; k < kp.K; k++) {
    c.a_ba[k] += c.P->Xs[(size_t)t * kp.K + k] * g_att;
    c.a_bd[k] += c.P->Xs[(size_t)t * kp.K + k] * g_def;
  }
  const double gx[4] = {g[eAh1], g[eAa1], -g[eBh1], -g[eBa1]};  // ha, aa, hd, ad
  if (kp.model == BPLX_DIXON_COLES) {
    c.a_mu[0] += gx[0];
  } else {
    const int nx = kp.model == BPLX_EXTENDED ? 1 : 4;
    for (int i = 0; i < nx; i++) {
      const double dec = c.th[o.dec[i] + t];
      c.grad[o.dec[i] + t] += c.sig[i] * gx[i];
      c.a_mu[i] += gx[i];
      c.a_ls[i] += c.sig[i] * dec * gx[i];
    }
  }
}

double sigmoid_clipped(double x, bool* clipped) {
  // numpyro SigmoidTransform: clip(expit(x), finfo.tiny, 1 - finfo.eps) -- in double here
  double s = 1.0 / (1.0 + std::exp(-x));
  *clipped = false;
  return s;
}

}  // namespace

extern "C" int bplx_plancheck_eval(const bplx_problem_desc* desc, const double* theta, double* lp_out, double* grad,
                                   double* corr_out, char* errbuf, int errlen) {
  HostPlan P;
  std::string err;
  int rc = build_plan(*desc, &P, &err);
  if (rc != BPLX_OK) {
    if (errbuf && errlen > 0) snprintf(errbuf, errlen, "%s", err.c_str());
    return rc;
  }
  const KernelParams& kp = P.kp;
  const ThetaOffsets& o = kp.off;
  const int T = kp.T, K = kp.K, V = kp.V, Cf = kp.Cf, D = kp.D;
  for (int i = 0; i < D; i++) grad[i] = 0.0;
  Ctx c;
  c.P = &P;
  c.kp = &kp;
  c.th = theta;
  c.grad = grad;
  c.a_conf.assign(Cf, 0.0);
  const bool dc = kp.model == BPLX_DIXON_COLES, ext = kp.model == BPLX_EXTENDED;
  const bool neu = kp.model == BPLX_NEUTRAL || kp.model == BPLX_NEUTRAL_WC;
  const bool has_rho = !dc;
  // ---- hypers -----------------------------------------------------------------------------------
  c.mu_d = theta[o.mean_defence];
  c.sig_a = std::exp(theta[o.log_std_attack]);
  c.sig_d = std::exp(theta[o.log_std_defence]);
  for (int i = 0; i < 4; i++) {
    c.mu[i] = o.mean[i] >= 0 ? theta[o.mean[i]] : 0.0;
    c.sig[i] = o.log_std[i] >= 0 ? std::exp(theta[o.log_std[i]]) : 0.0;
  }
  double u = 0.5;
  bool clipped;
  if (has_rho) u = sigmoid_clipped(theta[o.u], &clipped);
  c.rho = has_rho ? 2.0 * u - 1.0 : 0.0;
  c.s2 = 1.0 - c.rho * c.rho;
  for (int k = 0; k < K; k++) {
    c.beta_a[k] = theta[o.beta_a + k];
    c.beta_d[k] = theta[o.beta_d + k];
  }
  // ---- prologue: tables, team priors, static y part ---------------------------------------------
  const size_t rows = (size_t)V + 1;
  std::vector<Row> tab(kp.tab_bytes / kRowBytes + 4, Row{0, 0});  // one Row per table row
  auto row = [&](uint32_t off) -> Row& { return tab[off / kRowBytes]; };
  (void)rows;
  for (int t = 0; t < T; t++) {
    const double za = theta[o.za + t], zd = theta[o.zd + t];
    double am = 0, dm = c.mu_d;
    for (int k = 0; k < K; k++) {
      am += P.Xs[(size_t)t * K + k] * c.beta_a[k];
      dm += P.Xs[(size_t)t * K + k] * c.beta_d[k];
    }
    const double att = am + za * c.sig_a, def = dm + zd * c.sig_d;
    // priors on the standardised pair (extended_dixon_coles.py:165-174) or independent normals
    if (has_rho) {
      const double e = zd - c.rho * za;
      c.lp += -0.5 * za * za - kLogSqrt2Pi - 0.5 * e * e / c.s2 - 0.5 * std::log(c.s2) - kLogSqrt2Pi;
      grad[o.za + t] += -za + c.rho * e / c.s2;
      grad[o.zd + t] += -e / c.s2;
      c.a_rho += e * za / c.s2 - c.rho * e * e / (c.s2 * c.s2) + c.rho / c.s2;
    } else {
      c.lp += -0.5 * za * za - 0.5 * zd * zd - 2 * kLogSqrt2Pi;
      grad[o.za + t] += -za;
      grad[o.zd + t] += -zd;
    }
    double x[4] = {0, 0, 0, 0};
    if (dc) {
      x[0] = c.mu[0];
    } else {
      const int nx = ext ? 1 : 4;
      for (int i = 0; i < nx; i++) {
        const double dec = theta[o.dec[i] + t];
        x[i] = c.mu[i] + c.sig[i] * dec;
        c.lp += -0.5 * dec * dec - kLogSqrt2Pi;
        grad[o.dec[i] + t] += -dec;
      }
    }
    for (int v = P.team_vptr[t]; v < P.team_vptr[t + 1]; v++) {
      const double cf = kp.model == BPLX_NEUTRAL_WC ? theta[o.conf + P.v_conf[v]] : 0.0;
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
      double g[6];
      for (int e = 0; e < 6; e++) {
        g[e] = P.yexp[(size_t)v * 6 + e];
        c.lp += g[e] * ex[e];
      }
      apply_vteam(c, v, g);
    }
  }
  (void)neu;
  // ---- phase 1 ------------------------------------------------------------------------------------
  double best[3] = {0, 0, 0};  // max X, max Y, max XY over home-role lists
  int best_list[3] = {-1, -1, -1};
  double g[6];
  for (int w = 0; w < kp.nwarps; w++) {
    for (int li = P.warp_l1[w]; li < P.warp_l1[w + 1]; li++) {
      const List& L = P.lists1[li];
      if (L.flags & kListFirst) memset(g, 0, sizeof g);
      Row own = row(L.own_off);
      if (L.kind >= kH0) std::swap(own.x, own.y);
      const bool home = L.kind == kH1 || L.kind == kH0;
      double gx = 0, gy = 0;
      if (!kp.clip) {
        double ax = 0, ay = 0, m1 = 0, m2 = 0, m3 = 0;
        for (uint32_t i = 0; i < L.n; i++) {
          const Entry& e = P.ent1[L.ent + i];
          const Row& op = row(e.off);
          ax += e.w * op.x;
          ay += e.w * op.y;
          m1 = std::max(m1, op.x);
          m2 = std::max(m2, op.y);
          m3 = std::max(m3, op.x * op.y);
        }
        const double SX = own.x * ax, SY = own.y * ay;
        c.lp -= 0.5 * (SX + SY);
        gx = -SX;
        gy = -SY;
        if (home) {
          const double v[3] = {own.x * m1, own.y * m2, own.x * own.y * m3};
          for (int q = 0; q < 3; q++)
            if (v[q] > best[q]) best[q] = v[q], best_list[q] = li;
        }
      } else {
        for (uint32_t i = 0; i < L.n; i++) {
          const EntryClip& e = P.ent1c[L.ent + i];
          const Row& op = row(e.off);
          const double X = own.x * op.x, Y = own.y * op.y;
          const double Xc = std::min(X, 15.0), Yc = std::min(Y, 15.0);
          if (home) {
            c.lp += e.wyx * std::log(Xc) - e.w * Xc + e.wyy * std::log(Yc) - e.w * Yc;
            const double v[3] = {Xc, Yc, Xc * Yc};
            for (int q = 0; q < 3; q++)
              if (v[q] > best[q]) best[q] = v[q], best_list[q] = li;
          }
          if (X < 15.0) gx += e.wyx - e.w * X;
          if (Y < 15.0) gy += e.wyy - e.w * Y;
        }
      }
      g[kOwnX[L.kind]] += gx;
      g[kOwnY[L.kind]] += gy;
      if (L.flags & kListLast) apply_vteam(c, (int)L.vteam, g);
    }
  }
  // ---- bounds and corr_coef (bpl/_util.py:17-31) ---------------------------------------------------
  const double Lam = std::max(best[0], best[1]);
  const double LB = -1.0 / Lam;
  const double UB = std::min(1.0 / best[2], 1.0);
  const double r = sigmoid_clipped(theta[o.raw], &clipped);
  const double cc = LB + r * (UB - LB);
  *corr_out = cc;
  // ---- phase 2: tau terms ----------------------------------------------------------------------------
  double Gc = 0;
  for (int w = 0; w < kp.nwarps; w++) {
    for (int li = P.warp_l2[w]; li < P.warp_l2[w + 1]; li++) {
      const List& L = P.lists2[li];
      if (L.flags & kListFirst) memset(g, 0, sizeof g);
      Row own = row(L.own_off);
      if (L.kind >= kH0) std::swap(own.x, own.y);
      const bool home = L.kind == kH1 || L.kind == kH0;
      double uxy = 0, ux = 0, uy = 0, lt = 0;
      uint32_t i = L.ent;
      auto rates = [&](const Entry& e, double* X, double* Y) {
        const Row& op = row(e.off);
        *X = own.x * op.x;
        *Y = own.y * op.y;
        if (kp.clip) *X = std::min(*X, 15.0), *Y = std::min(*Y, 15.0);
      };
      double X, Y;
      for (uint32_t k = 0; k < L.n_xy; k++, i++) {
        const Entry& e = P.ent2[i];
        rates(e, &X, &Y);
        const double tau = 1.0 - cc * X * Y;
        uxy += e.w * X * Y / tau;
        if (e.w != 0) lt += e.w * std::log(tau);
      }
      for (uint32_t k = 0; k < L.n_x; k++, i++) {
        const Entry& e = P.ent2[i];
        rates(e, &X, &Y);
        const double tau = 1.0 + cc * X;
        ux += e.w * X / tau;
        if (e.w != 0) lt += e.w * std::log(tau);
      }
      for (uint32_t k = 0; k < L.n_y; k++, i++) {
        const Entry& e = P.ent2[i];
        rates(e, &X, &Y);
        const double tau = 1.0 + cc * Y;
        uy += e.w * Y / tau;
        if (e.w != 0) lt += e.w * std::log(tau);
      }
      // residuals wrt the two log-rates; a clipped rate has zero derivative
      double gx = cc * (ux - uxy), gy = cc * (uy - uxy);
      if (kp.clip) {
        // per-entry clip masks are needed: redo with masks (clip models are tiny; clarity over speed)
        gx = gy = 0;
        uint32_t j = L.ent;
        for (uint32_t k = 0; k < L.n_xy; k++, j++) {
          const Entry& e = P.ent2[j];
          const Row& op = row(e.off);
          const double Xr = own.x * op.x, Yr = own.y * op.y;
          rates(e, &X, &Y);
          const double q = e.w * cc * X * Y / (1.0 - cc * X * Y);
          if (Xr < 15.0) gx -= q;
          if (Yr < 15.0) gy -= q;
        }
        for (uint32_t k = 0; k < L.n_x; k++, j++) {
          const Entry& e = P.ent2[j];
          const Row& op = row(e.off);
          const double Xr = own.x * op.x;
          rates(e, &X, &Y);
          if (Xr < 15.0) gx += e.w * cc * X / (1.0 + cc * X);
        }
        for (uint32_t k = 0; k < L.n_y; k++, j++) {
          const Entry& e = P.ent2[j];
          const Row& op = row(e.off);
          const double Yr = own.y * op.y;
          rates(e, &X, &Y);
          if (Yr < 15.0) gy += e.w * cc * Y / (1.0 + cc * Y);
        }
      }
      if (home) {
        c.lp += lt;
        Gc += ux + uy - uxy;
      }
      g[kOwnX[L.kind]] += gx;
      g[kOwnY[L.kind]] += gy;
      if (L.flags & kListLast) apply_vteam(c, (int)L.vteam, g);
    }
  }
  c.lp += kp.w11 * std::log(1.0 - cc);
  Gc -= kp.w11 / (1.0 - cc);
  // ---- arg-max fix-up -----------------------------------------------------------
  // dc/dLB = 1 - r, dLB/dLam = 1/Lam^2, dLam/d eta = Lam  (zero when the max is a clipped rate)
  auto find = [&](int q, int* own_v, int* opp_v, int* kind, double* Xr, double* Yr) {
    const List& L = P.lists1[best_list[q]];
    Row own = row(L.own_off);
    if (L.kind >= kH0) std::swap(own.x, own.y);
    *own_v = (int)L.vteam;
    *kind = L.kind;
    *opp_v = -1;
    for (uint32_t i = 0; i < L.n; i++) {
      const uint32_t off = kp.clip ? P.ent1c[L.ent + i].off : P.ent1[L.ent + i].off;
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
      apply_vteam(c, ov, g1);
      apply_vteam(c, pv, g2);
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
    apply_vteam(c, ov, g1);
    apply_vteam(c, pv, g2);
  }
  // ---- hyper priors, Jacobians, chain rule -------------------------------------------------------------
  auto normal = [&](double x, double loc, double scale, int off) {
    const double z = (x - loc) / scale;
    c.lp += -0.5 * z * z - std::log(scale) - kLogSqrt2Pi;
    grad[off] += -z / scale;
  };
  auto halfnormal_exp = [&](double sig, double scale, int off, double acc) {
    // HalfNormal(scale) on sig = exp(x) plus the Jacobian x
    const double z = sig / scale;
    c.lp += -0.5 * z * z - std::log(scale) - kLogSqrt2Pi + kLog2 + theta[off];
    grad[off] += -z * z + 1.0 + acc;
  };
  const double std_scale = neu ? 0.5 : 1.0;  // neutral_dixon_coles.py:138-139
  normal(c.mu_d, 0.0, 1.0, o.mean_defence);
  grad[o.mean_defence] += c.a_mu_d;
  halfnormal_exp(c.sig_a, std_scale, o.log_std_attack, c.a_ls_a);
  halfnormal_exp(c.sig_d, std_scale, o.log_std_defence, c.a_ls_d);
  if (dc || ext) {
    normal(c.mu[0], 0.1, 0.2, o.mean[0]);
    grad[o.mean[0]] += c.a_mu[0];
    if (ext) halfnormal_exp(c.sig[0], 1.0, o.log_std[0], c.a_ls[0]);
  } else {
    const double loc[4] = {0.1, -0.1, 0.1, -0.1};
    for (int i = 0; i < 4; i++) {
      normal(c.mu[i], loc[i], 0.2, o.mean[i]);
      grad[o.mean[i]] += c.a_mu[i];
      halfnormal_exp(c.sig[i], 1.0, o.log_std[i], c.a_ls[i]);
    }
  }
  for (int k = 0; k < K; k++) {
    normal(c.beta_a[k], 0.0, 1.0, o.beta_a + k);
    grad[o.beta_a + k] += c.a_ba[k];
    normal(c.beta_d[k], 0.0, 1.0, o.beta_d + k);
    grad[o.beta_d + k] += c.a_bd[k];
  }
  if (has_rho) {  // u ~ Beta(2, 4) + sigmoid Jacobian; rho = 2u - 1
    c.lp += std::log(u) + 3.0 * std::log(1.0 - u) + std::log(20.0) + std::log(u) + std::log(1.0 - u);
    grad[o.u] += (1.0 - u) - 3.0 * u + (1.0 - 2.0 * u) + c.a_rho * 2.0 * u * (1.0 - u);
  }
  for (int k = 0; k < Cf; k++) {
    normal(theta[o.conf + k], 0.0, 1.0, o.conf + k);
    grad[o.conf + k] += c.a_conf[k];
  }
  // corr_coef_raw ~ Beta(2, 2) + Jacobian; c = LB + r (UB - LB)
  c.lp += 2.0 * (std::log(r) + std::log(1.0 - r)) + std::log(6.0);
  grad[o.raw] += 2.0 * (1.0 - 2.0 * r) + Gc * r * (1.0 - r) * (UB - LB);
  c.lp += kp.const_term;
  *lp_out = c.lp;
  return 0;
}
