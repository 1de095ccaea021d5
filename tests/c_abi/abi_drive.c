/* abi_drive.c -- plain C11 driver of include/bplx.h: create -> fwdbwd_host -> loglik layout -> score_grid_host -> destroy.
 *
 * Proves that the header is valid ISO C (built with -std=c11 -pedantic -Wall -Werror) and exercises the ABI without
 * ctypes / Python.  The data are a closed formula so that tests/test_c_abi.py can rebuild them and check the printed
 * numbers against the oracle.  Exit codes: 0 ok, 3 no CUDA device (bplx has no CPU fallback), 1 anything else. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "bplx.h"

#define T 5
#define M (T * (T - 1))
#define C 3

int main(void) {
  uint16_t home[M], away[M];
  uint8_t hg[M], ag[M];
  int m = 0, h, a, c, d;
  for (h = 0; h < T; h++)
    for (a = 0; a < T; a++) {
      if (h == a) continue;
      home[m] = (uint16_t)h;
      away[m] = (uint16_t)a;
      hg[m] = (uint8_t)((h * 3 + a * 5 + 1) % 4);
      ag[m] = (uint8_t)((h + 2 * a) % 3);
      m++;
    }
  bplx_problem_desc desc;
  desc.model = BPLX_DIXON_COLES;
  desc.num_matches = M;
  desc.num_teams = T;
  desc.num_covariates = 0;
  desc.num_conferences = 0;
  desc.num_gameweeks = 0;
  desc.flags = 0u;
  desc.home_team = home;
  desc.away_team = away;
  desc.home_goals = hg;
  desc.away_goals = ag;
  desc.neutral_venue = NULL;
  desc.home_conf = NULL;
  desc.away_conf = NULL;
  desc.gameweek = NULL;
  desc.weights = NULL;
  desc.covariates = NULL;

  bplx_problem* p = NULL;
  int rc = bplx_problem_create(&desc, &p);
  if (rc == BPLX_E_CUDA) {
    printf("nodevice: %s\n", bplx_last_error());
    return 3;
  }
  if (rc != BPLX_OK) {
    fprintf(stderr, "create failed (%d): %s\n", rc, bplx_last_error());
    return 1;
  }
  const int D = bplx_num_params(p);
  printf("version %d D %d layout %s\n", bplx_version(), D, bplx_problem_layout(p));
  printf("loglik_inputs %d loglik_layout %s\n", bplx_loglik_num_inputs(p), bplx_loglik_layout(p));
  float* theta = (float*)malloc(sizeof(float) * C * (size_t)D);
  float* grad = (float*)malloc(sizeof(float) * C * (size_t)D);
  float lp[C], cc[C];
  if (!theta || !grad) return 1;
  for (c = 0; c < C; c++)
    for (d = 0; d < D; d++) theta[c * D + d] = 0.3f * (float)sin(1.0 + d + 0.5 * c);
  rc = bplx_logdensity_fwdbwd_host(p, C, theta, lp, grad, cc);
  if (rc != BPLX_OK) {
    fprintf(stderr, "fwdbwd_host failed (%d): %s\n", rc, bplx_last_error());
    return 1;
  }
  for (c = 0; c < C; c++) {
    double n2 = 0.0;
    for (d = 0; d < D; d++) n2 += (double)grad[c * D + d] * grad[c * D + d];
    printf("chain %d lp %.9g corr_coef %.9g gradnorm %.9g\n", c, (double)lp[c], (double)cc[c], sqrt(n2));
  }
  /* predictive grid from two "posterior samples" (host arrays) */
  {
    float att[2 * T], def[2 * T], ha[2], corr[2] = {0.05f, -0.02f};
    uint16_t fh[2] = {0, 3}, fa[2] = {1, 2};
    float grid[2 * 6 * 6], outcome[2 * 3];
    int s, t;
    for (s = 0; s < 2; s++) {
      ha[s] = 0.2f + 0.1f * (float)s;
      for (t = 0; t < T; t++) {
        att[s * T + t] = 0.1f * (float)(t - 2) + 0.05f * (float)s;
        def[s * T + t] = 0.05f * (float)(2 - t);
      }
    }
    bplx_samples smp;
    smp.model = BPLX_DIXON_COLES;
    smp.num_samples = 2;
    smp.num_teams = T;
    smp.num_conferences = 0;
    smp.attack = att;
    smp.defence = def;
    smp.home_attack = ha;
    smp.away_attack = NULL;
    smp.home_defence = NULL;
    smp.away_defence = NULL;
    smp.confederation_strength = NULL;
    smp.corr_coef = corr;
    bplx_fixtures fx;
    fx.num_fixtures = 2;
    fx.home_team = fh;
    fx.away_team = fa;
    fx.home_conf = NULL;
    fx.away_conf = NULL;
    fx.neutral_venue = NULL;
    rc = bplx_score_grid_host(&smp, &fx, 5, 0.5f, grid, outcome);
    if (rc != BPLX_OK) {
      fprintf(stderr, "score_grid_host failed (%d): %s\n", rc, bplx_last_error());
      return 1;
    }
    printf("grid00 %.9g %.9g outcome0 %.9g %.9g %.9g\n", (double)grid[0], (double)grid[36], (double)outcome[0],
           (double)outcome[1], (double)outcome[2]);
    fh[1] = 77; /* out of range: must be refused, not read */
    rc = bplx_score_grid_host(&smp, &fx, 5, 0.5f, grid, outcome);
    printf("bad_fixture rc %d\n", rc);
  }
  free(theta);
  free(grad);
  bplx_problem_destroy(p);
  printf("launches %llu\n", bplx_launch_count());
  return 0;
}
