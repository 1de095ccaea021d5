// xla_ffi_shim.cc -- XLA FFI handlers over the C ABI (include/bplx.h), for registration with jax.ffi.
//
// NOT part of libbplx.so: it needs the XLA FFI headers that ship inside jaxlib (`jax.ffi.include_dir()`), and jax / jaxlib
// are not installable in this image.  Where they are:
//     g++ -O2 -std=c++17 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())") -I../../include \
//         -I/usr/local/cuda/include xla_ffi_shim.cc -L../lib -lbplx -o ../lib/libbplx_xla.so
// and INTEGRATION.md section B shows the Python side (jax.ffi.register_ffi_target + custom_vjp).  In this repository the
// file is compiled against a small stand-in for that API and its three implementations are called with device buffers
// (tests/xla_ffi_mock/, tests/test_xla_ffi_shim.py), so the argument unpacking below is exercised even without jaxlib.
// The handlers replace, for this path, what XLA compiles from the reference `_model`s (bpl/dixon_coles.py:39-84 etc.)
// inside numpyro's NUTS loop, and from the eager predict methods (bpl/base.py:74-148).
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime_api.h>

#include <cstdint>

#include "bplx.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

// leading axes of a [..., D] operand folded into one chain axis (jax.ffi.ffi_call(..., vmap_method="broadcast_all"))
template <typename B>
int64_t leading(const B& b) {
  const auto dims = b.dimensions();
  int64_t n = 1;
  for (size_t i = 0; i + 1 < dims.size(); i++) n *= dims[i];
  return n;
}

ffi::Error status(int rc) {
  return rc == BPLX_OK ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, bplx_last_error());
}

}  // namespace

// ---- full log-density (NUTS(potential_fn=...)) ---------------------------------------------------------------------
// theta [..., D] float32 -> lp [...], grad [..., D], corr_coef [...], workspace (uint8, bplx_logdensity_workspace_bytes)
ffi::Error LogDensityImpl(cudaStream_t stream, int64_t problem, ffi::Buffer<ffi::F32> theta,
                          ffi::ResultBuffer<ffi::F32> lp, ffi::ResultBuffer<ffi::F32> grad,
                          ffi::ResultBuffer<ffi::F32> corr_coef, ffi::ResultBuffer<ffi::U8> workspace) {
  const auto dims = theta.dimensions();
  if (dims.size() < 1) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "theta must have a trailing parameter axis");
  const bplx_problem* p = reinterpret_cast<const bplx_problem*>(problem);
  if (dims.back() != bplx_num_params(p))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "theta's last axis does not match the model's parameter count");
  return status(bplx_logdensity_fwdbwd(p, static_cast<int>(leading(theta)), BPLX_CHAIN_MAJOR, 0, theta.typed_data(),
                                       lp->typed_data(), grad->typed_data(), corr_coef->typed_data(),
                                       workspace->typed_data(), workspace->size_bytes(), stream));
}

// ---- likelihood only (numpyro.factor inside a `_model` that keeps its prior sites; dixon_coles.py:79-84) -------------
// tables [..., Dl] float32 packed per bplx_loglik_layout -> loglik [...], grad [..., Dl], corr_coef [...], workspace
ffi::Error LogLikImpl(cudaStream_t stream, int64_t problem, ffi::Buffer<ffi::F32> tables,
                      ffi::ResultBuffer<ffi::F32> loglik, ffi::ResultBuffer<ffi::F32> grad,
                      ffi::ResultBuffer<ffi::F32> corr_coef, ffi::ResultBuffer<ffi::U8> workspace) {
  const auto dims = tables.dimensions();
  const bplx_problem* p = reinterpret_cast<const bplx_problem*>(problem);
  if (dims.size() < 1 || dims.back() != bplx_loglik_num_inputs(p))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "tables' last axis does not match bplx_loglik_num_inputs");
  return status(bplx_loglik_fwdbwd(p, static_cast<int>(leading(tables)), BPLX_CHAIN_MAJOR, 0, tables.typed_data(),
                                   loglik->typed_data(), grad->typed_data(), corr_coef->typed_data(),
                                   workspace->typed_data(), workspace->size_bytes(), stream));
}

// ---- predictive grid (predict_score_grid_proba / predict_outcome_proba; base.py:74-148) -------------------------------
// posterior arrays [S, T] (home_advantage [S] for DIXON_COLES), corr_coef [S]; fixtures [F]; absent operands are passed
// as zero-sized buffers.  -> grid [F, g, g], outcome [F, 3], workspace (bplx_score_grid_workspace_bytes)
ffi::Error ScoreGridImpl(cudaStream_t stream, int64_t model, int64_t max_goals, float scale, ffi::Buffer<ffi::F32> attack,
                         ffi::Buffer<ffi::F32> defence, ffi::Buffer<ffi::F32> home_attack, ffi::Buffer<ffi::F32> away_attack,
                         ffi::Buffer<ffi::F32> home_defence, ffi::Buffer<ffi::F32> away_defence,
                         ffi::Buffer<ffi::F32> confederation_strength, ffi::Buffer<ffi::F32> corr_coef,
                         ffi::Buffer<ffi::U16> home_team, ffi::Buffer<ffi::U16> away_team, ffi::Buffer<ffi::U8> home_conf,
                         ffi::Buffer<ffi::U8> away_conf, ffi::Buffer<ffi::U8> neutral_venue,
                         ffi::ResultBuffer<ffi::F32> grid, ffi::ResultBuffer<ffi::F32> outcome,
                         ffi::ResultBuffer<ffi::U8> workspace) {
  const auto ad = attack.dimensions();
  if (ad.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "attack must be [S, T]");
  auto opt = [](auto& b) { return b.element_count() ? b.typed_data() : nullptr; };
  bplx_samples s;
  s.model = static_cast<int32_t>(model);
  s.num_samples = static_cast<int32_t>(ad[0]);
  s.num_teams = static_cast<int32_t>(ad[1]);
  s.num_conferences = confederation_strength.element_count() ? static_cast<int32_t>(confederation_strength.dimensions().back()) : 0;
  s.attack = attack.typed_data();
  s.defence = defence.typed_data();
  s.home_attack = opt(home_attack);
  s.away_attack = opt(away_attack);
  s.home_defence = opt(home_defence);
  s.away_defence = opt(away_defence);
  s.confederation_strength = opt(confederation_strength);
  s.corr_coef = corr_coef.typed_data();
  bplx_fixtures f;
  f.num_fixtures = static_cast<int32_t>(home_team.element_count());
  f.home_team = home_team.typed_data();
  f.away_team = away_team.typed_data();
  f.home_conf = opt(home_conf);
  f.away_conf = opt(away_conf);
  f.neutral_venue = opt(neutral_venue);
  return status(bplx_score_grid(&s, &f, static_cast<int>(max_goals), scale, grid->typed_data(), outcome->typed_data(),
                                workspace->typed_data(), workspace->size_bytes(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(bplx_logdensity_ffi, LogDensityImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("problem")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(bplx_loglik_ffi, LogLikImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("problem")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(bplx_score_grid_ffi, ScoreGridImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("model")
                                  .Attr<int64_t>("max_goals")
                                  .Attr<float>("scale")
                                  .Arg<ffi::Buffer<ffi::F32>>()   // attack
                                  .Arg<ffi::Buffer<ffi::F32>>()   // defence
                                  .Arg<ffi::Buffer<ffi::F32>>()   // home_attack (home_advantage for DC / EXTENDED)
                                  .Arg<ffi::Buffer<ffi::F32>>()   // away_attack
                                  .Arg<ffi::Buffer<ffi::F32>>()   // home_defence
                                  .Arg<ffi::Buffer<ffi::F32>>()   // away_defence
                                  .Arg<ffi::Buffer<ffi::F32>>()   // confederation_strength
                                  .Arg<ffi::Buffer<ffi::F32>>()   // corr_coef
                                  .Arg<ffi::Buffer<ffi::U16>>()   // home_team
                                  .Arg<ffi::Buffer<ffi::U16>>()   // away_team
                                  .Arg<ffi::Buffer<ffi::U8>>()    // home_conf
                                  .Arg<ffi::Buffer<ffi::U8>>()    // away_conf
                                  .Arg<ffi::Buffer<ffi::U8>>()    // neutral_venue
                                  .Ret<ffi::Buffer<ffi::F32>>()   // grid
                                  .Ret<ffi::Buffer<ffi::F32>>()   // outcome
                                  .Ret<ffi::Buffer<ffi::U8>>());  // workspace
#else
// XLA FFI headers absent: nothing to compile (the C ABI is driven through ctypes instead, bpl_next_b200/_abi.py).
#endif
