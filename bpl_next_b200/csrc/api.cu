// api.cu -- the extern "C" surface declared in include/bplx.h.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>

#include "common.cuh"
#include "plan.h"
#include "problem.h"
#include "score_grid.h"

namespace bplx {

static thread_local char g_err[1024] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
void note_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

static EnvSwitches g_env;
static std::once_flag g_env_once;
void reload_env_switches() {
  EnvSwitches e;
  e.no_pdl = getenv("BPLX_NO_PDL") != nullptr;
  e.nuts_generic = getenv("BPLX_NUTS_GENERIC") != nullptr;
  e.no_tail_split = getenv("BPLX_NO_TAIL_SPLIT") != nullptr;
  e.no_transpose = getenv("BPLX_NO_TRANSPOSE") != nullptr;
  if (const char* v = getenv("BPLX_TRANSPOSE_MIN_ELEMS")) e.transpose_min_elems = (size_t)atoll(v);
  if (const char* v = getenv("BPLX_SPLIT")) e.split = atoi(v);
  if (const char* v = getenv("BPLX_HOST_CHUNKS")) e.host_chunks = atoi(v);
  g_env = e;
}
const EnvSwitches& env_switches() {
  std::call_once(g_env_once, reload_env_switches);
  return g_env;
}

constexpr size_t kUploadSlack = (size_t)kMaxSplit * kMaxWarps * sizeof(EntryClip) + 64;
template <typename T>
static int upload(bplx_problem* p, const std::vector<T>& h, const T** out) {
  *out = nullptr;
  if (h.empty()) return BPLX_OK;
  void* d = nullptr;
  // kUploadSlack zero bytes follow every array: the arg-max search reads its first candidate entry of a piece before it
  // knows how many entries the piece has (logdensity.cu)
  const size_t bytes = h.size() * sizeof(T);
  BPLX_CUDA(cudaMalloc(&d, bytes + kUploadSlack));
  p->dev_allocs.push_back(d);
  BPLX_CUDA(cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice));
  BPLX_CUDA(cudaMemset(static_cast<unsigned char*>(d) + bytes, 0, kUploadSlack));
  *out = static_cast<const T*>(d);
  return BPLX_OK;
}

// the device-side alias of a page-locked (cudaHostAlloc / cudaHostRegister, mapped) host array, else `fallback`
static float* mapped_or(float* host, float* fallback) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
    cudaGetLastError();
    return fallback;
  }
  if (a.type == cudaMemoryTypeHost && a.devicePointer != nullptr) return static_cast<float*>(a.devicePointer);
  return fallback;
}

// out[c][r] = in[r][c] for an R x Cc matrix (row pitches ldi, ldo): the host entry point takes and returns [chains, D]
// arrays (what the reference's vmapped potential sees), the kernel's native layout is [D, chains].
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc,
                                                        size_t ldi, size_t ldo) {
  // a block moves 128 rows x 32 columns: its 16 loads per thread are all issued before the first store
  __shared__ float tile[4][32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 128;
  float v[4][4];
#pragma unroll
  for (int s = 0; s < 4; s++)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int r = r0 + s * 32 + threadIdx.y + 8 * k, c = c0 + threadIdx.x;
      v[s][k] = (r < R && c < Cc) ? in[(size_t)r * ldi + c] : 0.0f;
    }
#pragma unroll
  for (int s = 0; s < 4; s++)
#pragma unroll
    for (int k = 0; k < 4; k++) tile[s][threadIdx.y + 8 * k][threadIdx.x] = v[s][k];
  __syncthreads();
#pragma unroll
  for (int s = 0; s < 4; s++)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int c = c0 + threadIdx.y + 8 * k, r = r0 + s * 32 + threadIdx.x;
      if (r < R && c < Cc) out[(size_t)c * ldo + r] = tile[s][threadIdx.x][threadIdx.y + 8 * k];
    }
}
static int launch_transpose(const float* in, float* out, int R, int Cc, size_t ldi, size_t ldo, cudaStream_t s) {
  const dim3 grid((unsigned)((Cc + 31) / 32), (unsigned)((R + 127) / 128));
  BPLX_REQUIRE(grid.y <= 65535u, BPLX_E_UNSUPPORTED, "transpose of %d rows", R);
  transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(in, out, R, Cc, ldi, ldo);
  BPLX_CUDA(cudaGetLastError());
  note_launch(1);
  return BPLX_OK;
}

// Room behind the kernel's own workspace for one transposed copy of theta and of the gradient ([D][Cpad] each): large
// [chains, D] batches are computed in the kernel's native [D, chains] layout (see enqueue).
static size_t align256(size_t n) { return (n + 255) / 256 * 256; }
static size_t transpose_room(const KernelParams& kp, int C) {
  if ((size_t)C * (size_t)kp.D < env_switches().transpose_min_elems) return 0;
  return 2 * (((size_t)C + 31) / 32 * 32) * (size_t)kp.D * sizeof(float);
}

static size_t workspace_bytes(const KernelParams& kp, int C) {
  const size_t Cpad = ((size_t)C + 31) / 32 * 32;
  if (kp.model == BPLX_DYNAMIC)  // the walk's prefix sums (+ the per-gameweek hyper-parameter table when shared memory is short)
    return ((size_t)kp.G * kp.T * 2 + (kp.dyn_hyp_ws ? (size_t)kp.G * 16 : 0)) * Cpad * sizeof(float);
  if (kp.Cf <= 0) return 0;
  return (size_t)kp.V * Cpad * sizeof(float);
}

static int enqueue(const bplx_problem* p, int C, int layout, int ld, const float* theta, float* lp, float* grad,
                   float* corr_coef, void* ws, size_t ws_bytes, cudaStream_t stream, bool pdl = false, bool lik = false) {
  BPLX_REQUIRE(p != nullptr, BPLX_E_INVALID, "problem is NULL");
  BPLX_REQUIRE(C > 0, BPLX_E_INVALID, "num_chains must be positive (got %d)", C);
  BPLX_REQUIRE(theta && lp && grad, BPLX_E_INVALID, "theta, lp and grad must not be NULL");
  if (layout == BPLX_CHAIN_MAJOR && !env_switches().no_transpose) {
    // A [chains, D] batch (what a vmapped jax.ffi call hands over): tiled transposes of theta before and of the
    // gradient after the kernel, which then runs in its native layout (configs[2], 32,768 chains: 1.6 + 0.14 ms instead
    // of 3.5 ms of strided per-chain reads) -- when the caller's workspace has the room bplx_logdensity_workspace_bytes
    // asks for.  From 2^17 elements: measured per leapfrog of a fit with chain-major sampler state, configs[1] at 4,096
    // chains x 73 parameters 45 -> 40 us, configs[2] data at 256 chains x 1,339 231 -> 153 us; configs[0] at 1,024 x 45
    // (below the threshold) 28.7 vs 29.5 us.
    const size_t base = align256(workspace_bytes(p->kp, C)), room = transpose_room(p->kp, C);
    if (room > 0 && ws != nullptr && ws_bytes >= base + room) {
      const int Dc = lik ? p->lik_D : p->kp.D;
      const size_t ldm = ld > 0 ? (size_t)ld : (size_t)Dc, Cpad = ((size_t)C + 31) / 32 * 32;
      BPLX_REQUIRE(ldm >= (size_t)Dc, BPLX_E_INVALID, "chain-major ld (%d) < D (%d)", ld, Dc);
      float* tin = reinterpret_cast<float*>(static_cast<char*>(ws) + base);
      float* tout = tin + Cpad * (size_t)p->kp.D;
      int rc = launch_transpose(theta, tin, C, Dc, ldm, Cpad, stream);
      if (rc != BPLX_OK) return rc;
      rc = enqueue(p, C, BPLX_CHAIN_MINOR, (int)Cpad, tin, lp, tout, corr_coef, ws, base, stream, pdl, lik);
      if (rc != BPLX_OK) return rc;
      return launch_transpose(tout, grad, Dc, C, Cpad, ldm, stream);
    }
  }
  KernelParams kp = p->kp;
  if (lik) {  // the same plan read through the likelihood-only site layout: constrained tables in, no priors
    BPLX_REQUIRE(kp.model != BPLX_DYNAMIC, BPLX_E_UNSUPPORTED, "bplx_loglik_fwdbwd: DYNAMIC is not supported");
    kp.lik_only = 1;
    kp.off = p->lik_off;
    for (int i = 0; i < 12; i++) kp.hyper[i] = p->lik_hyper[i];
    kp.nhyper = p->lik_nhyper;
    kp.D = p->lik_D;
    kp.K = 0;
    kp.const_term = p->lik_const;
  }
  if (layout == BPLX_CHAIN_MAJOR) {
    kp.sd = 1;
    kp.sc = ld > 0 ? ld : kp.D;
    BPLX_REQUIRE(kp.sc >= kp.D, BPLX_E_INVALID, "chain-major ld (%d) < D (%d)", ld, kp.D);
  } else if (layout == BPLX_CHAIN_MINOR) {
    kp.sc = 1;
    kp.sd = ld > 0 ? ld : C;
    BPLX_REQUIRE(kp.sd >= C, BPLX_E_INVALID, "chain-minor ld (%d) < num_chains (%d)", ld, C);
  } else {
    BPLX_REQUIRE(false, BPLX_E_INVALID, "unknown layout %d", layout);
  }
  BPLX_REQUIRE((long long)kp.D * (long long)kp.sd + (long long)C * (long long)kp.sc < (1ll << 31), BPLX_E_UNSUPPORTED,
               "theta / grad of %d x %d elements exceed the 31-bit element index of the kernel", C, kp.D);
  const size_t need = workspace_bytes(kp, C);
  BPLX_REQUIRE(ws_bytes >= need && (need == 0 || ws != nullptr), BPLX_E_WORKSPACE,
               "workspace too small: %zu bytes given, %zu needed", ws_bytes, need);
  kp.C = C;
  kp.Cpad = (C + 31) / 32 * 32;
  kp.theta = theta;
  kp.lp = lp;
  kp.grad = grad;
  kp.corr_coef = corr_coef;
  kp.scratch = static_cast<float*>(ws);
  // few chains: several CTAs (one cluster) share a group of 32 chains, as many as the device holds at once
  int si = 0;
  const int groups = (C + kChains - 1) / kChains;
  for (int i = kNumSplits - 1; i >= 1; i--)
    if (p->s1[i] && (1 << i) <= kp.split_hint && groups <= p->max_clusters[i]) {
      si = i;
      break;
    }
  if (const int want = env_switches().split) {  // testing / tuning: force 1, 2, 4 or 8 CTAs per group
    for (int i = 0; i < kNumSplits; i++)
      if ((1 << i) == want && p->s1[i] && (i == 0 || p->max_clusters[i] > 0)) si = i;
  }
  auto use_split = [&](int i) {
    kp.split = 1 << i;
    kp.stream1 = p->s1[i];
    kp.stream2 = p->s2[i];
    kp.warp_b1 = p->wb1[i];
    kp.warp_b2 = p->wb2[i];
  };
  use_split(si);
  kp.group0 = 0;
  kp.ngroups = groups;
  if (kp.model == BPLX_DYNAMIC) return launch_logdensity_dynamic(kp, stream);
  // Tail split: the kernel holds one CTA per SM, so `groups` CTAs run in waves of `sm_count`; when the last, partial
  // wave is small enough, its groups go to a second launch with 2, 4 or 8 CTAs (a cluster) per group, which fills the
  // SMs the wave would leave idle and ends sooner (configs[2] at 16,384 chains: 512 groups = 3 waves + 68 groups).
  int ti = 0;
  const int rest = p->sm_count > 0 ? groups % p->sm_count : 0;
  if (si == 0 && rest > 0 && rest < groups && !env_switches().split && !env_switches().no_tail_split) {
    for (int i = kNumSplits - 1; i >= 1; i--)
      if (p->s1[i] && (1 << i) <= kp.split_hint && rest <= p->max_clusters[i] && rest * (1 << i) <= p->sm_count) {
        ti = i;
        break;
      }
  }
  if (ti == 0) return launch_logdensity(kp, p->wb[si], stream, pdl);
  kp.ngroups = groups - rest;
  int rc = launch_logdensity(kp, p->wb[0], stream, pdl);
  if (rc != BPLX_OK) return rc;
  use_split(ti);
  kp.group0 = groups - rest;
  kp.ngroups = rest;
  return launch_logdensity(kp, p->wb[ti], stream, pdl);
}

}  // namespace bplx

using namespace bplx;

extern "C" {

int bplx_problem_create(const bplx_problem_desc* desc, bplx_problem** out) {
  BPLX_REQUIRE(desc && out, BPLX_E_INVALID, "desc / out is NULL");
  *out = nullptr;
  HostPlan hp;
  std::string err;
  int rc = build_plan(*desc, &hp, &err);
  if (rc != BPLX_OK) {
    set_error("%s", err.c_str());
    return rc;
  }
  bplx_problem* p = new (std::nothrow) bplx_problem();
  BPLX_REQUIRE(p, BPLX_E_NOMEM, "out of host memory");
  auto fail = [&](int code) {
    bplx_problem_destroy(p);
    return code;
  };
  if (cudaGetDevice(&p->device) != cudaSuccess) {
    set_error("cudaGetDevice failed: no usable CUDA device (bplx has no CPU fallback)");
    return fail(BPLX_E_CUDA);
  }
  p->kp = hp.kp;
  p->layout = hp.layout;
  p->lik_off = hp.lik_off;
  for (int i = 0; i < 12; i++) p->lik_hyper[i] = hp.lik_hyper[i];
  p->lik_nhyper = hp.lik_nhyper;
  p->lik_D = hp.lik_D;
  p->lik_const = hp.lik_const;
  p->lik_layout = hp.lik_layout;
  KernelParams& kp = p->kp;
#define UP(vec, ptr)                                 \
  if ((rc = upload(p, hp.vec, &(ptr))) != BPLX_OK) return fail(rc)
  UP(stream1, p->s1[0]);
  UP(stream2, p->s2[0]);
  UP(warp_b1, p->wb1[0]);
  UP(warp_b2, p->wb2[0]);
  for (int i = 1; i < kNumSplits; i++) {
    if (hp.more[i - 1].stream1.empty()) continue;  // (DYNAMIC has no split plans)
    UP(more[i - 1].stream1, p->s1[i]);
    UP(more[i - 1].stream2, p->s2[i]);
    UP(more[i - 1].warp_b1, p->wb1[i]);
    UP(more[i - 1].warp_b2, p->wb2[i]);
  }
  auto keep_bounds = [&](int i, const std::vector<uint32_t>& b1, const std::vector<uint32_t>& b2) {
    const size_t cap = sizeof(p->wb[i].b1) / sizeof(uint32_t);
    for (size_t k = 0; k < b1.size() && k < cap; k++) p->wb[i].b1[k] = b1[k];
    for (size_t k = 0; k < b2.size() && k < cap; k++) p->wb[i].b2[k] = b2[k];
  };
  keep_bounds(0, hp.warp_b1, hp.warp_b2);
  for (int i = 1; i < kNumSplits; i++) keep_bounds(i, hp.more[i - 1].warp_b1, hp.more[i - 1].warp_b2);
  kp.stream1 = p->s1[0];
  kp.stream2 = p->s2[0];
  kp.warp_b1 = p->wb1[0];
  kp.warp_b2 = p->wb2[0];
  kp.split = 1;
  UP(team_vptr, kp.team_vptr);
  UP(team_flags, kp.team_flags);
  UP(v_team, kp.v_team);
  UP(v_conf, kp.v_conf);
  UP(conf_vptr, kp.conf_vptr);
  UP(conf_vlist, kp.conf_vlist);
  UP(yexp, kp.yexp);
  UP(yteam, kp.yteam);
  UP(yconf, kp.yconf);
  UP(Xs, kp.Xs);
  UP(gw_tptr, kp.gw_tptr);
  UP(gw_tlist, kp.gw_tlist);
#undef UP
  rc = kp.model == BPLX_DYNAMIC ? logdensity_dynamic_set_attributes() : logdensity_set_attributes(kp);
  if (rc != BPLX_OK) return fail(rc);
  if (cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, p->device) != cudaSuccess) p->sm_count = 0;
  p->max_clusters[0] = 1 << 30;
  if (kp.model != BPLX_DYNAMIC)
    for (int i = 1; i < kNumSplits; i++) p->max_clusters[i] = p->s1[i] ? logdensity_max_clusters(kp, 1 << i) : 0;
  p->stats[0] = desc->num_matches;
  p->stats[1] = hp.n1;
  p->stats[2] = hp.n1_padded;
  p->stats[3] = hp.n2;
  p->stats[4] = hp.n2_padded;
  p->stats[5] = kp.smem_total;
  p->stats[6] = kp.nwarps;
  p->stats[7] = kp.V;
  if (kp.model != BPLX_DYNAMIC) {  // what each warp of the split-1 plan walks (plan tuning)
    p->warp_stats.assign((size_t)kp.nwarps * 12, 0);
    for (int phase = 0; phase < 2; phase++) {
      const std::vector<unsigned char>& S = phase ? hp.stream2 : hp.stream1;
      const std::vector<uint32_t>& wb = phase ? hp.warp_b2 : hp.warp_b1;
      const size_t esz = (phase == 0 && kp.clip) ? sizeof(EntryClip) : sizeof(Entry);
      const uint32_t minp = phase ? kp.min_piece2 : kp.min_piece1;
      for (int w = 0; w < kp.nwarps && w + 1 < (int)wb.size(); w++) {
        long long* o = &p->warp_stats[(size_t)w * 12 + phase * 6];
        const uint32_t len = wb[w + 1] - wb[w];
        o[5] = (len + kp.stage_bytes - 1) / kp.stage_bytes;
        for (uint32_t st = 0; st < len; st += kp.stage_bytes) {
          uint32_t a = wb[w] + st;
          const uint32_t aend = wb[w] + std::min(len, st + kp.stage_bytes);
          while (a + minp <= aend) {
            ListHdr H;
            memcpy(&H, &S[a], sizeof H);
            if (H.n0 + H.n1 + H.n2 == 0) break;  // stage padding
            const bool home = (H.kind & 1) == 0;
            if (phase == 0) {
              o[home ? 0 : 1]++;
              o[home ? 2 : 3] += H.n0;
            } else {
              o[0]++;
              o[1] += H.n0;
              o[2] += H.n1;
              o[3] += H.n2;
            }
            if (H.flags & kTeamFirst) o[4]++;
            a += (uint32_t)sizeof(ListHdr) + (uint32_t)((H.n0 + H.n1 + H.n2) * esz);
          }
        }
      }
    }
  }
  *out = p;
  return BPLX_OK;
}

void bplx_problem_destroy(bplx_problem* p) {
  if (!p) return;
  for (void* d : p->dev_allocs) cudaFree(d);
  if (p->d_theta) cudaFree(p->d_theta);
  if (p->d_out) cudaFree(p->d_out);
  if (p->d_ws) cudaFree(p->d_ws);
  if (p->d_tin) cudaFree(p->d_tin);
  if (p->d_tout) cudaFree(p->d_tout);
  if (p->host_stream) cudaStreamDestroy(p->host_stream);
  if (p->host_in) cudaStreamDestroy(p->host_in);
  if (p->host_out) cudaStreamDestroy(p->host_out);
  for (cudaEvent_t e : p->host_ev)
    if (e) cudaEventDestroy(e);
  delete p;
}

int bplx_num_params(const bplx_problem* p) { return p ? p->kp.D : BPLX_E_INVALID; }

const char* bplx_problem_layout(const bplx_problem* p) { return p ? p->layout.c_str() : ""; }

int bplx_problem_stats(const bplx_problem* p, long long* out, int n) {
  BPLX_REQUIRE(p && out, BPLX_E_INVALID, "problem / out is NULL");
  for (int i = 0; i < n && i < 8; i++) out[i] = p->stats[i];
  return n < 8 ? n : 8;
}

int bplx_problem_warp_stats(const bplx_problem* p, long long* out, int n) {
  BPLX_REQUIRE(p && out, BPLX_E_INVALID, "problem / out is NULL");
  const int have = (int)p->warp_stats.size();
  for (int i = 0; i < n && i < have; i++) out[i] = p->warp_stats[i];
  return n < have ? n : have;
}

size_t bplx_logdensity_workspace_bytes(const bplx_problem* p, int num_chains) {
  if (!p || num_chains <= 0) return 0;
  const size_t room = transpose_room(p->kp, num_chains);
  return room ? align256(workspace_bytes(p->kp, num_chains)) + room : workspace_bytes(p->kp, num_chains);
}

int bplx_logdensity_fwdbwd(const bplx_problem* p, int num_chains, int layout, int ld, const float* theta, float* lp,
                           float* grad, float* corr_coef, void* workspace, size_t workspace_bytes, void* stream) {
  return enqueue(p, num_chains, layout, ld, theta, lp, grad, corr_coef, workspace, workspace_bytes,
                 static_cast<cudaStream_t>(stream), /*pdl=*/true);
}

int bplx_loglik_num_inputs(const bplx_problem* p) {
  if (!p) return BPLX_E_INVALID;
  return p->kp.model == BPLX_DYNAMIC ? BPLX_E_UNSUPPORTED : p->lik_D;
}

const char* bplx_loglik_layout(const bplx_problem* p) { return p ? p->lik_layout.c_str() : ""; }

int bplx_loglik_fwdbwd(const bplx_problem* p, int num_chains, int layout, int ld, const float* tables, float* loglik,
                       float* grad, float* corr_coef, void* workspace, size_t workspace_bytes, void* stream) {
  return enqueue(p, num_chains, layout, ld, tables, loglik, grad, corr_coef, workspace, workspace_bytes,
                 static_cast<cudaStream_t>(stream), /*pdl=*/true, /*lik=*/true);
}

int bplx_logdensity_fwdbwd_host(bplx_problem* p, int C, const float* theta, float* lp, float* grad, float* corr_coef) {
  BPLX_REQUIRE(p != nullptr, BPLX_E_INVALID, "problem is NULL");
  BPLX_REQUIRE(C > 0, BPLX_E_INVALID, "num_chains must be positive (got %d)", C);
  BPLX_REQUIRE(theta && lp && grad, BPLX_E_INVALID, "theta, lp and grad must not be NULL");
  std::lock_guard<std::mutex> lock(p->mu);
  BPLX_CUDA(cudaSetDevice(p->device));
  const size_t D = (size_t)p->kp.D;
  if (!p->host_stream) BPLX_CUDA(cudaStreamCreateWithFlags(&p->host_stream, cudaStreamNonBlocking));
  if (C > p->host_cap) {
    if (p->d_theta) cudaFree(p->d_theta);
    if (p->d_out) cudaFree(p->d_out);
    if (p->d_ws) cudaFree(p->d_ws);
    p->d_theta = p->d_out = nullptr;
    p->d_ws = nullptr;
    p->host_cap = 0;
    BPLX_CUDA(cudaMalloc(&p->d_theta, (size_t)C * D * sizeof(float)));
    BPLX_CUDA(cudaMalloc(&p->d_out, ((size_t)C * D + 2 * (size_t)C) * sizeof(float)));
    p->d_ws_bytes = workspace_bytes(p->kp, C);
    if (p->d_ws_bytes) BPLX_CUDA(cudaMalloc(&p->d_ws, p->d_ws_bytes));
    p->host_cap = C;
  }
  // Chains are cut into up to four chunks so that the upload of chunk k+1, the kernel of chunk k and the download
  // of chunk k-1 overlap (PCIe is full duplex); the kernels run in order on one stream and share the workspace.
  if (!p->host_in) {
    BPLX_CUDA(cudaStreamCreateWithFlags(&p->host_in, cudaStreamNonBlocking));
    BPLX_CUDA(cudaStreamCreateWithFlags(&p->host_out, cudaStreamNonBlocking));
    for (cudaEvent_t& e : p->host_ev) BPLX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  cudaStream_t s = p->host_stream;
  float* d_grad = p->d_out;
  float* d_lp = p->d_out + (size_t)C * D;
  float* d_cc = d_lp + C;
  // (only worth it when a chunk still fills the GPU: below ~16k chains K1 is latency-bound and chunking serialises it;
  //  measured on configs[1], 4,096 chains: 104 us in one piece, 138 us in four)
  // by payload (measured on configs[2] data, ms per call for 1 / 2 / 4 / 8 chunks: 22 MB each way 1.07 / 0.82 / 0.78 / 1.01,
  // 44 MB 2.10 / 1.53 / 1.32 / 1.41, 88 MB 4.08 / 2.97 / 2.48 / 2.39; 175 MB: 8 chunks)
  const size_t payload = (size_t)C * D * sizeof(float);
  int nchunk = payload >= ((size_t)80 << 20) ? 8 : (payload >= ((size_t)16 << 20) ? 4 : (payload >= ((size_t)6 << 20) ? 2 : 1));
  while (nchunk > 1 && C / nchunk < 256) nchunk /= 2;
  if (const int want = env_switches().host_chunks)  // tuning: 1, 2, 4, 8 or 16 pipelined chunks
    if (want == 1 || want == 2 || want == 4 || want == 8 || want == 16) nchunk = want;
  // Large batches go through the kernel's native chain-minor layout: a tiled transpose of theta before the kernel and of
  // the gradient after it (two passes over HBM each, ~0.1 ms for configs[2] at 32,768 chains) instead of the kernel's
  // strided per-chain reads (3.5 ms against 1.6 ms in the native layout).
  const bool native = (size_t)C * D >= std::max(env_switches().transpose_min_elems, (size_t)1 << 20) && !env_switches().no_transpose;
  if (nchunk == 1) {  // one piece: everything in order on one stream, no events (each costs microseconds at this scale)
    // lp and corr_coef are one coalesced 128-byte store per warp: when the caller's arrays are page-locked the kernel
    // writes them straight into host memory (posted PCIe writes) instead of two more copies of 6.6 us each
    float* k_lp = mapped_or(lp, d_lp);
    float* k_cc = corr_coef ? mapped_or(corr_coef, d_cc) : d_cc;
    BPLX_CUDA(cudaMemcpyAsync(p->d_theta, theta, (size_t)C * D * sizeof(float), cudaMemcpyHostToDevice, s));
    int rc;
    if (native) {
      const int ldc = (C + 31) / 32 * 32;
      if (ldc > p->tr_cap) {
        if (p->d_tin) cudaFree(p->d_tin);
        if (p->d_tout) cudaFree(p->d_tout);
        p->d_tin = p->d_tout = nullptr;
        p->tr_cap = 0;
        BPLX_CUDA(cudaMalloc(&p->d_tin, (size_t)ldc * D * sizeof(float)));
        BPLX_CUDA(cudaMalloc(&p->d_tout, (size_t)ldc * D * sizeof(float)));
        p->tr_cap = ldc;
      }
      rc = launch_transpose(p->d_theta, p->d_tin, C, (int)D, D, (size_t)ldc, s);
      if (rc != BPLX_OK) return rc;
      rc = enqueue(p, C, BPLX_CHAIN_MINOR, ldc, p->d_tin, k_lp, p->d_tout, k_cc, p->d_ws, p->d_ws_bytes, s);
      if (rc != BPLX_OK) return rc;
      rc = launch_transpose(p->d_tout, d_grad, (int)D, C, (size_t)ldc, D, s);
    } else {
      rc = enqueue(p, C, BPLX_CHAIN_MAJOR, (int)D, p->d_theta, k_lp, d_grad, k_cc, p->d_ws, p->d_ws_bytes, s);
    }
    if (rc != BPLX_OK) return rc;
    BPLX_CUDA(cudaMemcpyAsync(grad, d_grad, (size_t)C * D * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (k_lp == d_lp) BPLX_CUDA(cudaMemcpyAsync(lp, d_lp, (size_t)C * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (corr_coef && k_cc == d_cc)
      BPLX_CUDA(cudaMemcpyAsync(corr_coef, d_cc, (size_t)C * sizeof(float), cudaMemcpyDeviceToHost, s));
    BPLX_CUDA(cudaStreamSynchronize(s));
    return BPLX_OK;
  }
  const int per = ((C + nchunk - 1) / nchunk + 31) / 32 * 32;
  if (native && per > p->tr_cap) {
    if (p->d_tin) cudaFree(p->d_tin);
    if (p->d_tout) cudaFree(p->d_tout);
    p->d_tin = p->d_tout = nullptr;
    p->tr_cap = 0;
    BPLX_CUDA(cudaMalloc(&p->d_tin, (size_t)per * D * sizeof(float)));
    BPLX_CUDA(cudaMalloc(&p->d_tout, (size_t)per * D * sizeof(float)));
    p->tr_cap = per;
  }
  // chunk sizes: with four or more chunks the first and the last are cut in half -- while the first chunk uploads and the
  // last downloads, the link runs in one direction only
  int sizes[40], ns = 0, left = C;
  auto push = [&](int n) {
    n = n < left ? n : left;
    if (n > 0) sizes[ns++] = n, left -= n;
  };
  if (nchunk >= 4) {
    const int half = (per / 2 + 31) / 32 * 32;
    push(half);
    push(half);
    while (left > per) push(per);
    push((left / 2 + 31) / 32 * 32);
    push(left);
  } else {
    while (left > 0) push(per);
  }
  for (int k = 0, c0 = 0; k < ns; c0 += sizes[k], k++) {
    const int n = sizes[k];
    const size_t off = (size_t)c0 * D;
    BPLX_CUDA(cudaMemcpyAsync(p->d_theta + off, theta + off, (size_t)n * D * sizeof(float), cudaMemcpyHostToDevice, p->host_in));
    BPLX_CUDA(cudaEventRecord(p->host_ev[2 * k], p->host_in));
    BPLX_CUDA(cudaStreamWaitEvent(s, p->host_ev[2 * k], 0));
    int rc;
    if (native) {  // (the transposed buffers are shared by the chunks: everything that touches them is in order on `s`)
      rc = launch_transpose(p->d_theta + off, p->d_tin, n, (int)D, D, (size_t)per, s);
      if (rc != BPLX_OK) return rc;
      rc = enqueue(p, n, BPLX_CHAIN_MINOR, per, p->d_tin, d_lp + c0, p->d_tout, d_cc + c0, p->d_ws, p->d_ws_bytes, s);
      if (rc != BPLX_OK) return rc;
      rc = launch_transpose(p->d_tout, d_grad + off, (int)D, n, (size_t)per, D, s);
    } else {
      rc = enqueue(p, n, BPLX_CHAIN_MAJOR, (int)D, p->d_theta + off, d_lp + c0, d_grad + off, d_cc + c0, p->d_ws,
                   p->d_ws_bytes, s);
    }
    if (rc != BPLX_OK) return rc;
    BPLX_CUDA(cudaEventRecord(p->host_ev[2 * k + 1], s));
    BPLX_CUDA(cudaStreamWaitEvent(p->host_out, p->host_ev[2 * k + 1], 0));
    BPLX_CUDA(cudaMemcpyAsync(grad + off, d_grad + off, (size_t)n * D * sizeof(float), cudaMemcpyDeviceToHost, p->host_out));
    BPLX_CUDA(cudaMemcpyAsync(lp + c0, d_lp + c0, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, p->host_out));
    if (corr_coef)
      BPLX_CUDA(cudaMemcpyAsync(corr_coef + c0, d_cc + c0, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, p->host_out));
  }
  BPLX_CUDA(cudaStreamSynchronize(p->host_out));
  return BPLX_OK;
}

// ---- score grid ------------------------------------------------------------------------------------------
static int check_grid_args(const bplx_samples* s, const bplx_fixtures* f, int max_goals) {
  BPLX_REQUIRE(s && f, BPLX_E_INVALID, "samples / fixtures is NULL");
  BPLX_REQUIRE(s->model >= BPLX_DIXON_COLES && s->model <= BPLX_NEUTRAL_WC, BPLX_E_UNSUPPORTED,
               "score grid: model %d unsupported (DYNAMIC predict is unusable in the reference)", s->model);
  BPLX_REQUIRE(s->num_samples > 0 && s->num_teams > 0 && f->num_fixtures > 0, BPLX_E_INVALID,
               "score grid: num_samples, num_teams and num_fixtures must be positive");
  BPLX_REQUIRE(max_goals >= 1 && max_goals <= 63, BPLX_E_INVALID, "max_goals must be in [1, 63] (got %d)", max_goals);
  BPLX_REQUIRE(s->attack && s->defence && s->corr_coef && f->home_team && f->away_team, BPLX_E_INVALID,
               "score grid: attack, defence, corr_coef, home_team, away_team must not be NULL");
  BPLX_REQUIRE(s->home_attack, BPLX_E_INVALID, "score grid: home_attack (home advantage) must not be NULL");
  if (s->model == BPLX_NEUTRAL || s->model == BPLX_NEUTRAL_WC)
    BPLX_REQUIRE(s->away_attack && s->home_defence && s->away_defence, BPLX_E_INVALID,
                 "score grid: neutral models need away_attack, home_defence, away_defence");
  if (s->model == BPLX_NEUTRAL_WC)
    BPLX_REQUIRE(s->confederation_strength && s->num_conferences > 0 && f->home_conf && f->away_conf, BPLX_E_INVALID,
                 "score grid: NEUTRAL_WC needs confederation_strength, home_conf, away_conf");
  return BPLX_OK;
}

static size_t grid_plan(const bplx_samples* s, const bplx_fixtures* f, int max_goals, GridParams* gp, const char** err) {
  gp->model = s->model;
  gp->S = s->num_samples;
  gp->T = s->num_teams;
  gp->Cf = s->model == BPLX_NEUTRAL_WC ? s->num_conferences : 0;
  gp->F = f->num_fixtures;
  gp->g = max_goals + 1;
  return score_grid_plan(gp, err);
}

size_t bplx_score_grid_workspace_bytes(const bplx_samples* s, const bplx_fixtures* f, int max_goals) {
  if (!s || !f || max_goals < 1 || s->num_samples <= 0 || s->num_teams <= 0 || f->num_fixtures <= 0) return 0;
  GridParams gp{};
  return grid_plan(s, f, max_goals, &gp, nullptr);
}

int bplx_score_grid(const bplx_samples* s, const bplx_fixtures* f, int max_goals, float scale, float* grid,
                    float* outcome, void* workspace, size_t workspace_bytes, void* stream) {
  return bplx_score_grid_ex(s, f, max_goals, scale, grid, outcome, workspace, workspace_bytes, stream, 0u);
}

int bplx_score_grid_ex(const bplx_samples* s, const bplx_fixtures* f, int max_goals, float scale, float* grid,
                       float* outcome, void* workspace, size_t workspace_bytes, void* stream, unsigned flags) {
  int rc = check_grid_args(s, f, max_goals);
  if (rc != BPLX_OK) return rc;
  BPLX_REQUIRE(grid != nullptr, BPLX_E_INVALID, "grid is NULL");
  GridParams gp{};
  const char* perr = nullptr;
  const size_t need = grid_plan(s, f, max_goals, &gp, &perr);
  BPLX_REQUIRE(need > 0, BPLX_E_UNSUPPORTED, "%s", perr ? perr : "score grid: unsupported shape");
  gp.scale = scale;
  BPLX_REQUIRE(workspace && workspace_bytes >= need, BPLX_E_WORKSPACE,
               "workspace too small: %zu bytes given, %zu needed", workspace_bytes, need);
  BPLX_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, BPLX_E_INVALID, "workspace must be 16-byte aligned");
  gp.attack = s->attack;
  gp.defence = s->defence;
  gp.ha = s->home_attack;
  gp.aa = s->away_attack;
  gp.hd = s->home_defence;
  gp.ad = s->away_defence;
  gp.conf = s->confederation_strength;
  gp.corr = s->corr_coef;
  gp.home = f->home_team;
  gp.away = f->away_team;
  gp.hconf = f->home_conf;
  gp.aconf = f->away_conf;
  gp.nv = f->neutral_venue;
  gp.table = static_cast<float*>(workspace);
  gp.partial = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) +
                                        ((size_t)gp.S * gp.row_floats * 4 + 255) / 256 * 256);
  gp.grid = grid;
  gp.outcome = outcome;
  gp.reuse_tables = (flags & BPLX_GRID_REUSE_TABLES) ? 1 : 0;
  return launch_score_grid(gp, static_cast<cudaStream_t>(stream));
}

int bplx_score_grid_host(const bplx_samples* s, const bplx_fixtures* f, int max_goals, float scale, float* grid,
                         float* outcome) {
  int rc = check_grid_args(s, f, max_goals);
  if (rc != BPLX_OK) return rc;
  BPLX_REQUIRE(grid != nullptr, BPLX_E_INVALID, "grid is NULL");
  const size_t S = s->num_samples, T = s->num_teams, F = f->num_fixtures, g = max_goals + 1;
  const size_t Cf = s->model == BPLX_NEUTRAL_WC ? s->num_conferences : 0;
  for (size_t i = 0; i < F; i++) {  // host arrays: an out-of-range index would read past the staged sample rows
    BPLX_REQUIRE(f->home_team[i] < T && f->away_team[i] < T, BPLX_E_INVALID,
                 "score grid: team index out of range at fixture %zu (%u, %u; num_teams %zu)", i,
                 (unsigned)f->home_team[i], (unsigned)f->away_team[i], T);
    if (Cf)
      BPLX_REQUIRE(f->home_conf[i] < Cf && f->away_conf[i] < Cf, BPLX_E_INVALID,
                   "score grid: confederation index out of range at fixture %zu (%u, %u; num_conferences %zu)", i,
                   (unsigned)f->home_conf[i], (unsigned)f->away_conf[i], Cf);
  }
  std::vector<void*> allocs;
  auto cleanup = [&]() {
    for (void* d : allocs) cudaFree(d);
  };
  cudaStream_t st = nullptr;
  if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("cudaStreamCreate failed: no usable CUDA device (bplx has no CPU fallback)");
    return BPLX_E_CUDA;
  }
  rc = BPLX_OK;
  auto up = [&](const void* h, size_t bytes) -> void* {
    if (!h || rc != BPLX_OK) return nullptr;
    void* d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess ||
        cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) {
      set_error("score_grid_host: device allocation / copy of %zu bytes failed", bytes);
      rc = BPLX_E_CUDA;
      return nullptr;
    }
    allocs.push_back(d);
    return d;
  };
  bplx_samples ds = *s;
  bplx_fixtures df = *f;
  const size_t ha_w = s->model == BPLX_DIXON_COLES ? 1 : T;
  ds.attack = (const float*)up(s->attack, S * T * 4);
  ds.defence = (const float*)up(s->defence, S * T * 4);
  ds.home_attack = (const float*)up(s->home_attack, S * ha_w * 4);
  const bool neu = s->model == BPLX_NEUTRAL || s->model == BPLX_NEUTRAL_WC;
  ds.away_attack = neu ? (const float*)up(s->away_attack, S * T * 4) : nullptr;
  ds.home_defence = neu ? (const float*)up(s->home_defence, S * T * 4) : nullptr;
  ds.away_defence = neu ? (const float*)up(s->away_defence, S * T * 4) : nullptr;
  ds.confederation_strength = Cf ? (const float*)up(s->confederation_strength, S * Cf * 4) : nullptr;
  ds.corr_coef = (const float*)up(s->corr_coef, S * 4);
  df.home_team = (const uint16_t*)up(f->home_team, F * 2);
  df.away_team = (const uint16_t*)up(f->away_team, F * 2);
  df.home_conf = Cf ? (const uint8_t*)up(f->home_conf, F) : nullptr;
  df.away_conf = Cf ? (const uint8_t*)up(f->away_conf, F) : nullptr;
  df.neutral_venue = (neu && f->neutral_venue) ? (const uint8_t*)up(f->neutral_venue, F) : nullptr;
  const size_t ws_bytes = bplx_score_grid_workspace_bytes(s, f, max_goals);
  void *d_ws = nullptr, *d_grid = nullptr, *d_out = nullptr;
  if (rc == BPLX_OK) {
    if (cudaMalloc(&d_ws, ws_bytes) == cudaSuccess) allocs.push_back(d_ws); else rc = BPLX_E_NOMEM;
    if (cudaMalloc(&d_grid, F * g * g * 4) == cudaSuccess) allocs.push_back(d_grid); else rc = BPLX_E_NOMEM;
    if (outcome) {
      if (cudaMalloc(&d_out, F * 3 * 4) == cudaSuccess) allocs.push_back(d_out); else rc = BPLX_E_NOMEM;
    }
    if (rc != BPLX_OK) set_error("score_grid_host: out of device memory");
  }
  if (rc == BPLX_OK)
    rc = bplx_score_grid(&ds, &df, max_goals, scale, (float*)d_grid, (float*)d_out, d_ws, ws_bytes, st);
  if (rc == BPLX_OK) {
    cudaError_t e = cudaMemcpyAsync(grid, d_grid, F * g * g * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && outcome) e = cudaMemcpyAsync(outcome, d_out, F * 3 * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      set_error("score_grid_host: %s", cudaGetErrorString(e));
      rc = BPLX_E_CUDA;
    }
  } else {
    cudaStreamSynchronize(st);
  }
  cleanup();
  cudaStreamDestroy(st);
  return rc;
}

void bplx_reload_env(void) {
  env_switches();
  reload_env_switches();
}
const char* bplx_last_error(void) { return g_err; }
int bplx_version(void) { return BPLX_VERSION; }
unsigned long long bplx_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
