// TEST SCAFFOLDING: calls the implementations of bpl_next_b200/csrc/xla_ffi_shim.cc (compiled against the stand-in API next
// to this file) with device pointers handed over from Python -- what XLA's executor would do with the operand buffers.
#include "../../bpl_next_b200/csrc/xla_ffi_shim.cc"

#include <cstring>

namespace {
int finish(const ffi::Error& e, char* err, int errlen) {
  if (!e.failure()) return 0;
  if (err && errlen > 0) {
    std::strncpy(err, e.message().c_str(), (size_t)errlen - 1);
    err[errlen - 1] = 0;
  }
  return -1;
}
}  // namespace

extern "C" int mock_xla_call_density(int lik, int64_t problem, float* x, int64_t batch0, int64_t batch1, int64_t D, float* lp,
                                     float* grad, float* cc, void* ws, int64_t ws_bytes, void* stream, char* err, int errlen) {
  // a [batch0, batch1, D] operand: the handler must fold the two leading axes into one chain axis
  ffi::Buffer<ffi::F32> in(x, {batch0, batch1, D});
  ffi::ResultBuffer<ffi::F32> o_lp(ffi::Buffer<ffi::F32>(lp, {batch0, batch1}));
  ffi::ResultBuffer<ffi::F32> o_g(ffi::Buffer<ffi::F32>(grad, {batch0, batch1, D}));
  ffi::ResultBuffer<ffi::F32> o_cc(ffi::Buffer<ffi::F32>(cc, {batch0, batch1}));
  ffi::ResultBuffer<ffi::U8> o_ws(ffi::Buffer<ffi::U8>(ws, {ws_bytes}));
  auto* fn = lik ? bplx_loglik_ffi : bplx_logdensity_ffi;
  return finish(fn(static_cast<cudaStream_t>(stream), problem, in, o_lp, o_g, o_cc, o_ws), err, errlen);
}

extern "C" int mock_xla_call_grid(int64_t model, int64_t max_goals, float scale, int64_t S, int64_t T, int64_t Cf, int64_t F,
                                  float* attack, float* defence, float* ha, float* aa, float* hd, float* ad, float* conf,
                                  float* corr, uint16_t* home, uint16_t* away, uint8_t* hconf, uint8_t* aconf, uint8_t* nv,
                                  int64_t ha_cols, float* grid, float* outcome, void* ws, int64_t ws_bytes, void* stream,
                                  char* err, int errlen) {
  const int64_t g = max_goals + 1;
  auto f32 = [&](float* p, int64_t rows, int64_t cols) {
    return p ? ffi::Buffer<ffi::F32>(p, {rows, cols}) : ffi::Buffer<ffi::F32>(nullptr, {0});
  };
  auto u8 = [&](uint8_t* p) { return p ? ffi::Buffer<ffi::U8>(p, {F}) : ffi::Buffer<ffi::U8>(nullptr, {0}); };
  ffi::ResultBuffer<ffi::F32> o_grid(ffi::Buffer<ffi::F32>(grid, {F, g, g}));
  ffi::ResultBuffer<ffi::F32> o_out(ffi::Buffer<ffi::F32>(outcome, {F, 3}));
  ffi::ResultBuffer<ffi::U8> o_ws(ffi::Buffer<ffi::U8>(ws, {ws_bytes}));
  return finish(bplx_score_grid_ffi(static_cast<cudaStream_t>(stream), model, max_goals, scale, f32(attack, S, T),
                                    f32(defence, S, T), f32(ha, S, ha_cols), f32(aa, S, T), f32(hd, S, T), f32(ad, S, T),
                                    f32(conf, S, Cf), ffi::Buffer<ffi::F32>(corr, {S}), ffi::Buffer<ffi::U16>(home, {F}),
                                    ffi::Buffer<ffi::U16>(away, {F}), u8(hconf), u8(aconf), u8(nv), o_grid, o_out, o_ws),
                err, errlen);
}
