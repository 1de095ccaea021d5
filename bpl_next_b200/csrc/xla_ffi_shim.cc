// xla_ffi_shim.cc -- XLA FFI handlers over the C ABI (include/bplx.h), for registration with jax.ffi.
//
// NOT part of libbplx.so and not built in this image: it needs the XLA FFI headers that ship inside jaxlib
// (`jax.ffi.include_dir()`), and jax / jaxlib are not installed here.  Where they are:
//     g++ -O2 -std=c++17 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())") -I../../include \
//         xla_ffi_shim.cc -L../lib -lbplx -o ../lib/libbplx_xla.so
// and INTEGRATION.md section B shows the Python side (jax.ffi.register_ffi_target + custom_vjp).  It replaces, for this
// path, what XLA compiles from the reference `_model`s (bpl/dixon_coles.py:39-84 etc.) inside numpyro's NUTS loop.
#if __has_include("xla/ffi/api/ffi.h")
#include <cstdint>

#include "bplx.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

// theta [..., D] float32 (chain-major, batch dimensions folded into C by vmap_method="broadcast_all")
// -> lp [...], grad [..., D], corr_coef [...], workspace (uint8, bplx_logdensity_workspace_bytes(problem, C) bytes)
static ffi::Error LogDensityImpl(cudaStream_t stream, int64_t problem, ffi::Buffer<ffi::F32> theta,
                                 ffi::ResultBuffer<ffi::F32> lp, ffi::ResultBuffer<ffi::F32> grad,
                                 ffi::ResultBuffer<ffi::F32> corr_coef, ffi::ResultBuffer<ffi::U8> workspace) {
  const auto dims = theta.dimensions();
  if (dims.size() < 1) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "theta must have a trailing parameter axis");
  int64_t chains = 1;
  for (size_t i = 0; i + 1 < dims.size(); i++) chains *= dims[i];
  const bplx_problem* p = reinterpret_cast<const bplx_problem*>(problem);
  if (dims.back() != bplx_num_params(p))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "theta's last axis does not match the model's parameter count");
  const int rc = bplx_logdensity_fwdbwd(p, static_cast<int>(chains), BPLX_CHAIN_MAJOR, 0, theta.typed_data(),
                                        lp->typed_data(), grad->typed_data(), corr_coef->typed_data(),
                                        workspace->typed_data(), workspace->size_bytes(), stream);
  return rc == BPLX_OK ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, bplx_last_error());
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
#else
// XLA FFI headers absent: nothing to compile (the C ABI is driven through ctypes instead, bpl_next_b200/_abi.py).
#endif
