// score_grid.cu -- K3: posterior-predictive score grid + outcome probabilities.
//
// Replaces predict_score_grid_proba / predict_outcome_proba (bpl/base.py:74-148,
// bpl/neutral_dixon_coles.py:562-659, bpl/neutral_dixon_coles_WC.py:548-670) and, inside them,
// _calculate_expected_goals + predict_score_proba (bpl/dixon_coles.py:126-163,
// bpl/extended_dixon_coles.py:335-399, bpl/neutral_dixon_coles.py:399-488,
// bpl/neutral_dixon_coles_WC.py:363-474) and dixon_coles_correlation_term with weights=None
// (bpl/_util.py:35-93).  The reference materialises [S, F*g*g] temporaries and recomputes the two
// rates g*g times per fixture; here one thread owns one fixture, keeps its g x g tile of the grid in
// registers, and walks the posterior samples, which are staged [samples x teams] in shared memory
// with cp.async double buffering.  Poisson pmfs come from the recurrence p_k = p_{k-1} lambda / k.
//
//   grid = (F / 256 fixture blocks, grid tiles, sample splits); partial sums per split go to the
//   workspace and a finalize kernel sums the splits in a fixed order (deterministic), applies
//   `scale` and the 1 / (hg! ag!) of the two pmfs, and a third kernel reduces the W/D/L masks of
//   bpl/base.py:140-142.
#include "score_grid.h"

namespace bplx {

namespace {

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// cooperative copy of `n` floats (all threads of the CTA)
__device__ __forceinline__ void stage_copy(float* dst, const float* src, int n) {
  const uint32_t d = smem_u32(dst);
  if ((((uintptr_t)src | (uintptr_t)d) & 15) == 0 && (n & 3) == 0) {
    for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4) cp_async16(d + i * 4, src + i);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) cp_async4(d + i * 4, src + i);
  }
}

struct StageLayout {
  int att, def, ha, aa, hd, ad, conf, corr, total;  // float offsets inside one stage
  int ha_w;                                         // row width of `ha` (1 for DIXON_COLES)
};

__host__ __device__ inline StageLayout stage_layout(int model, int T, int Cf) {
  StageLayout L{};
  const int SB = kGridStage;
  int o = 0;
  L.att = o; o += SB * T;
  L.def = o; o += SB * T;
  L.ha_w = model == BPLX_DIXON_COLES ? 1 : T;
  L.ha = o; o += (SB * L.ha_w + 3) / 4 * 4;
  L.aa = L.hd = L.ad = L.conf = 0;
  if (model == BPLX_NEUTRAL || model == BPLX_NEUTRAL_WC) {
    L.aa = o; o += SB * T;
    L.hd = o; o += SB * T;
    L.ad = o; o += SB * T;
  }
  if (model == BPLX_NEUTRAL_WC) {
    L.conf = o; o += (SB * Cf + 3) / 4 * 4;
  }
  L.corr = o; o += SB;
  L.total = (o + 3) / 4 * 4;
  return L;
}

__device__ __forceinline__ void load_stage(const GridParams& gp, const StageLayout& L, float* buf, int s0, int ns) {
  const size_t T = gp.T;
  stage_copy(buf + L.att, gp.attack + (size_t)s0 * T, ns * gp.T);
  stage_copy(buf + L.def, gp.defence + (size_t)s0 * T, ns * gp.T);
  stage_copy(buf + L.ha, gp.ha + (size_t)s0 * L.ha_w, ns * L.ha_w);
  if (gp.model == BPLX_NEUTRAL || gp.model == BPLX_NEUTRAL_WC) {
    stage_copy(buf + L.aa, gp.aa + (size_t)s0 * T, ns * gp.T);
    stage_copy(buf + L.hd, gp.hd + (size_t)s0 * T, ns * gp.T);
    stage_copy(buf + L.ad, gp.ad + (size_t)s0 * T, ns * gp.T);
  }
  if (gp.model == BPLX_NEUTRAL_WC) stage_copy(buf + L.conf, gp.conf + (size_t)s0 * gp.Cf, ns * gp.Cf);
  stage_copy(buf + L.corr, gp.corr + s0, ns);
}

}  // namespace

// R x CC tile of the grid per thread; SINGLE = the tile is the whole grid (row/col origin 0).
template <int R, int CC, bool SINGLE>
__global__ void __launch_bounds__(kGridThreads, 1) score_grid_kernel(const __grid_constant__ GridParams gp) {
  extern __shared__ __align__(16) float sbuf[];
  const StageLayout L = stage_layout(gp.model, gp.T, gp.Cf);
  const int f_raw = blockIdx.x * kGridThreads + threadIdx.x;
  const int f = min(f_raw, gp.F - 1);
  const int tiles_c = (gp.g + CC - 1) / CC;
  const int r0 = SINGLE ? 0 : (blockIdx.y / tiles_c) * R;
  const int c0 = SINGLE ? 0 : (blockIdx.y % tiles_c) * CC;
  const bool corner = SINGLE || (r0 == 0 && c0 == 0);
  const int split = blockIdx.z;
  const int s_begin = split * gp.samples_per_split;
  const int s_end = min(gp.S, s_begin + gp.samples_per_split);

  const int h = gp.home[f], a = gp.away[f];
  const bool neu = gp.model == BPLX_NEUTRAL || gp.model == BPLX_NEUTRAL_WC;
  const float n = (neu && gp.nv && gp.nv[f]) ? 0.0f : 1.0f;  // 1 - neutral_venue
  const int hc = gp.model == BPLX_NEUTRAL_WC ? gp.hconf[f] : 0;
  const int ac = gp.model == BPLX_NEUTRAL_WC ? gp.aconf[f] : 0;

  float acc[R][CC];
#pragma unroll
  for (int i = 0; i < R; i++)
#pragma unroll
    for (int j = 0; j < CC; j++) acc[i][j] = 0.0f;

  const int nst = (s_end - s_begin + kGridStage - 1) / kGridStage;
  if (nst > 0) load_stage(gp, L, sbuf, s_begin, min(kGridStage, s_end - s_begin));
  cp_async_commit();
  for (int st = 0; st < nst; st++) {
    float* cur = sbuf + (st & 1) * L.total;
    if (st + 1 < nst) {
      const int s1 = s_begin + (st + 1) * kGridStage;
      load_stage(gp, L, sbuf + ((st + 1) & 1) * L.total, s1, min(kGridStage, s_end - s1));
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int ns = min(kGridStage, s_end - (s_begin + st * kGridStage));
    // Software pipeline over the samples of the stage: the rates and pmf powers of sample j+1 (gathers -> exp -> exp ->
    // recurrence, one long dependent chain) are computed in the same basic block as the R x CC FMAs of sample j, so the
    // scheduler hides the chain behind the FMA stream (two warps per scheduler cannot hide it by themselves).
    // The accumulators hold sum_s e^-lh lh^i e^-la la^j; 1 / (i! j!) is applied once, by the finalize kernel.
    auto vectors = [&](int j, bool live, float (&ph)[R], float (&pa)[CC], float (&tt)[4]) {
      const float* A = cur + L.att + j * gp.T;
      const float* Dd = cur + L.def + j * gp.T;
      float eh = A[h] - Dd[a], ea = A[a] - Dd[h];
      if (gp.model == BPLX_DIXON_COLES) {
        eh += cur[L.ha + j];
      } else if (gp.model == BPLX_EXTENDED) {
        eh += cur[L.ha + j * gp.T + h];
      } else {
        if (gp.model == BPLX_NEUTRAL_WC) {
          const float* cs = cur + L.conf + j * gp.Cf;
          const float dcf = cs[hc] - cs[ac];
          eh += dcf;
          ea -= dcf;
        }
        eh += n * cur[L.ha + j * gp.T + h] - n * cur[L.ad + j * gp.T + a];
        ea += n * cur[L.aa + j * gp.T + a] - n * cur[L.hd + j * gp.T + h];
      }
      const float lh = __expf(eh), la = __expf(ea);
      const float c = cur[L.corr + j];
      float p = live ? __expf(-lh) : 0.0f;  // a sample past the end of the stage gets weight 0
      if (!SINGLE)
        for (int k = 1; k <= r0; k++) p *= lh;
      ph[0] = p;
#pragma unroll
      for (int i = 1; i < R; i++) {
        p *= lh;
        ph[i] = p;
      }
      p = __expf(-la);
      if (!SINGLE)
        for (int k = 1; k <= c0; k++) p *= la;
      pa[0] = p;
#pragma unroll
      for (int i = 1; i < CC; i++) {
        p *= la;
        pa[i] = p;
      }
      // tau on the four low-score cells, clipped at 0 (bpl/_util.py:62-68)
      tt[0] = tt[1] = tt[2] = tt[3] = 1.0f;
      if (corner) {
        tt[0] = fmaxf(1.0f - (c * lh) * la, 0.0f);
        tt[1] = fmaxf(fmaf(c, lh, 1.0f), 0.0f);  // home 0, away 1
        tt[2] = fmaxf(fmaf(c, la, 1.0f), 0.0f);  // home 1, away 0
        tt[3] = fmaxf(1.0f - c, 0.0f);
      }
    };
    float ph[R], pa[CC], tt[4];
    vectors(0, true, ph, pa, tt);
#pragma unroll 1
    for (int j = 0; j < ns; j++) {
      float ph_n[R], pa_n[CC], tt_n[4];
      vectors(min(j + 1, ns - 1), j + 1 < ns, ph_n, pa_n, tt_n);
#pragma unroll
      for (int i = 0; i < R; i++) {
#pragma unroll
        for (int jj = 0; jj < CC; jj++) {
          float w = ph[i];
          if (i < 2 && jj < 2) w *= tt[i * 2 + jj];
          acc[i][jj] = fmaf(w, pa[jj], acc[i][jj]);
        }
      }
#pragma unroll
      for (int i = 0; i < R; i++) ph[i] = ph_n[i];
#pragma unroll
      for (int i = 0; i < CC; i++) pa[i] = pa_n[i];
#pragma unroll
      for (int i = 0; i < 4; i++) tt[i] = tt_n[i];
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  if (f_raw < gp.F) {
    float* out = gp.partial + ((size_t)split * gp.F + f) * (size_t)(gp.g * gp.g);
#pragma unroll
    for (int i = 0; i < R; i++)
#pragma unroll
      for (int jj = 0; jj < CC; jj++)
        if (r0 + i < gp.g && c0 + jj < gp.g) out[(r0 + i) * gp.g + c0 + jj] = acc[i][jj];
  }
}

// sums the sample splits in a fixed order, applies scale; one thread per grid element
__global__ void score_grid_finalize(const GridParams gp) {
  const size_t n = (size_t)gp.F * gp.g * gp.g;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.0f;
  for (int k = 0; k < gp.nsplit; k++) s += gp.partial[(size_t)k * n + i];
  // the kernel accumulates e^-lh lh^i e^-la la^j: the pmfs' 1 / (i! j!) goes here (double: 63! overflows float)
  const int cell = (int)(i % (size_t)(gp.g * gp.g)), hg = cell / gp.g, ag = cell % gp.g;
  double f = 1.0;
  for (int k = 2; k <= hg; k++) f *= (double)k;
  for (int k = 2; k <= ag; k++) f *= (double)k;
  gp.grid[i] = (float)((double)s * (double)gp.scale / f);
}

// home_win / draw / away_win = masked sums of the grid (bpl/base.py:140-142); one warp per fixture
__global__ void score_grid_outcome(const GridParams gp) {
  const int f = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (f >= gp.F) return;
  const int g = gp.g;
  const float* p = gp.grid + (size_t)f * g * g;
  float hw = 0.0f, dr = 0.0f, aw = 0.0f;
  for (int i = lane; i < g * g; i += 32) {
    const int hg = i / g, ag = i % g;
    const float v = p[i];
    if (hg > ag) hw += v;
    else if (hg == ag) dr += v;
    else aw += v;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    hw += __shfl_xor_sync(0xffffffffu, hw, o);
    dr += __shfl_xor_sync(0xffffffffu, dr, o);
    aw += __shfl_xor_sync(0xffffffffu, aw, o);
  }
  if (lane == 0) {
    gp.outcome[(size_t)f * 3 + 0] = hw;
    gp.outcome[(size_t)f * 3 + 1] = dr;
    gp.outcome[(size_t)f * 3 + 2] = aw;
  }
}

// ---- host side ---------------------------------------------------------------------------------------
static void tile_shape(int g, int* R, int* CC, int* ntiles) {
  if (g <= 11) {
    *R = 11; *CC = 11; *ntiles = 1;
  } else {
    *R = 8; *CC = 16;
    *ntiles = ((g + 7) / 8) * ((g + 15) / 16);
  }
}

size_t score_grid_workspace(int S, int F, int g, int* nsplit, int* samples_per_split) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  int R, CC, ntiles;
  tile_shape(g, &R, &CC, &ntiles);
  const int nfb = (F + kGridThreads - 1) / kGridThreads;
  const int max_split = S / 64 > 1 ? (S / 64 < 32 ? S / 64 : 32) : 1;
  int best = 1;
  double best_eff = -1.0;
  for (int ns = 1; ns <= max_split; ns++) {
    const long long ctas = (long long)nfb * ntiles * ns;
    const long long waves = (ctas + sms - 1) / sms;
    const double eff = (double)ctas / (double)(waves * sms);
    if (eff > best_eff + 1e-9) best_eff = eff, best = ns;
  }
  int sps = (S + best - 1) / best;
  sps = (sps + kGridStage - 1) / kGridStage * kGridStage;
  best = (S + sps - 1) / sps;
  if (nsplit) *nsplit = best;
  if (samples_per_split) *samples_per_split = sps;
  return (size_t)best * F * g * g * sizeof(float);
}

template <int R, int CC, bool SINGLE>
static int launch_tile(const GridParams& gp, int ntiles, size_t smem, cudaStream_t stream) {
  auto* fn = &score_grid_kernel<R, CC, SINGLE>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    BPLX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  dim3 grid((gp.F + kGridThreads - 1) / kGridThreads, ntiles, gp.nsplit);
  fn<<<grid, kGridThreads, smem, stream>>>(gp);
  BPLX_CUDA(cudaGetLastError());
  note_launch(1);
  return BPLX_OK;
}

int launch_score_grid(const GridParams& gp, cudaStream_t stream) {
  const StageLayout L = stage_layout(gp.model, gp.T, gp.Cf);
  const size_t smem = (size_t)2 * L.total * sizeof(float);
  BPLX_REQUIRE(smem <= 227 * 1024, BPLX_E_UNSUPPORTED,
               "score grid needs %zu bytes of shared memory per CTA (max %d): too many teams (%d)", smem, 227 * 1024,
               gp.T);
  int R, CC, ntiles;
  tile_shape(gp.g, &R, &CC, &ntiles);
  int rc = gp.g <= 11 ? launch_tile<11, 11, true>(gp, ntiles, smem, stream)
                      : launch_tile<8, 16, false>(gp, ntiles, smem, stream);
  if (rc != BPLX_OK) return rc;
  const size_t n = (size_t)gp.F * gp.g * gp.g;
  score_grid_finalize<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(gp);
  BPLX_CUDA(cudaGetLastError());
  note_launch(1);
  if (gp.outcome) {
    score_grid_outcome<<<(gp.F + 7) / 8, 256, 0, stream>>>(gp);
    BPLX_CUDA(cudaGetLastError());
    note_launch(1);
  }
  return BPLX_OK;
}

}  // namespace bplx
