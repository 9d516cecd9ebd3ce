// Test harness (TEST INFRASTRUCTURE): compiles integration/xla_ffi/eincm_xla_ffi.cc UNCHANGED against the stand-in FFI header
// (tests/native/mock_xla_ffi) and exposes its two handler bodies behind a C interface, so that pytest can drive them with device
// pointers the way XLA would: eager set_window per window, then evaluations that only carry theta and the token.
#include "../../integration/xla_ffi/eincm_xla_ffi.cc"

#include <cstdio>

namespace {
int report(const ffi::Error& e, char* msg, int msg_len) {
    if (msg != nullptr && msg_len > 0) std::snprintf(msg, (size_t)msg_len, "%s", e.message().c_str());
    return e.code();
}
}  // namespace

extern "C" {

int ffi_set_window(void* stream, int device, int16_t* xs, int16_t* ys, double* ts, long long n, double* edges, int R, int H, int W, double* edge_ts_dev,
                   int slot, int64_t* token_out_dev, char* msg, int msg_len) {
    ffi::ResultBuffer<ffi::S64> token(ffi::Buffer<ffi::S64>(token_out_dev, {2}));
    return report(SetWindowImpl((cudaStream_t)stream, device, ffi::Buffer<ffi::S16>(xs, {n}), ffi::Buffer<ffi::S16>(ys, {n}), ffi::Buffer<ffi::F64>(ts, {n}),
                                ffi::Buffer<ffi::F64>(edges, {R, H, W}), ffi::Buffer<ffi::F64>(edge_ts_dev, {R}), slot, token), msg, msg_len);
}

int ffi_value_and_grad(void* stream, int device, double* theta, int h, int w, int64_t* token_dev, double alpha, double beta, double gamma, double delta,
                       int cur_pyr_lvl, int n_pyr_lvls, int slot, double* loss_out_dev, double* grad_out_dev, char* msg, int msg_len) {
    ffi::ResultBuffer<ffi::F64> loss(ffi::Buffer<ffi::F64>(loss_out_dev, {})), grad(ffi::Buffer<ffi::F64>(grad_out_dev, {h, w, 2}));
    return report(ValueAndGradImpl((cudaStream_t)stream, device, ffi::Buffer<ffi::F64>(theta, {h, w, 2}), ffi::Buffer<ffi::S64>(token_dev, {2}), alpha, beta,
                                   gamma, delta, cur_pyr_lvl, n_pyr_lvls, slot, loss, grad), msg, msg_len);
}

// the binder expressions type-check against the stand-in as well
int ffi_binders_ok() { return EincmSetWindow_mock_bound() && EincmValueAndGrad_mock_bound(); }

}  // extern "C"
