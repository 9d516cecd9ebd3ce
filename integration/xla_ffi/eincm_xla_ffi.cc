// XLA FFI handler that registers the library's objective + gradient as a JAX custom call on the CUDA platform.
//
// NOT COMPILED OR TESTED IN THIS REPOSITORY'S IMAGE: it needs the XLA FFI headers that ship with jaxlib
// (`python -c "import jax; print(jax.ffi.include_dir())"`), and jax / jaxlib are not installed here (no network).  It is written
// against the public `xla/ffi/api/ffi.h` C++ API (jaxlib >= 0.4.31) and only calls entry points declared in include/eincm.h, all of
// which are tested through ctypes (tests/).  The handler body has been type-checked with g++ against a minimal mock of the FFI types;
// the binder expression at the end of the file has not.  Build where jaxlib exists:
//
//   g++ -O2 -std=c++17 -shared -fPIC eincm_xla_ffi.cc -o libeincm_xla_ffi.so
//       -I"$(python -c 'import jax; print(jax.ffi.include_dir())')" -I../../include -I/usr/local/cuda/include
//       -L../../edge-informed-contrast-maximization_b200/lib -leincm_b200 -Wl,-rpath,'$ORIGIN'        (one command line)
//
// Python side: integration/xla_ffi/eincm_jax.py.
//
// What it replaces: the body of `loss_func` (reference src/eincm/losses.py:108-205) under `jit(value_and_grad(...))` as jaxopt's
// ScipyMinimize builds it (reference src/eincm/solver.py:165-183).  XLA owns the buffers and the stream; the handler only enqueues.
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include <vector>

#include <cuda_runtime_api.h>

#include "eincm.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

// One plan per (device, H, W, event capacity, reference capacity).  XLA may call the handler from any host thread: the cache and
// every plan are guarded by one mutex (plans are thread-compatible, not thread-safe; an evaluation is a handful of launches, so the
// critical section is microseconds).
struct Entry {
    eincm_plan* plan = nullptr;
    int64_t window_id = -1;             // the window currently staged in the plan
};
std::mutex g_mu;
std::map<std::tuple<int, int, int, int64_t, int>, Entry> g_plans;

Entry* entry_for(int device, int H, int W, int64_t n_events, int R, const char** err) {
    // capacities are rounded up so that windows of slightly different sizes share a plan
    const int64_t cap = ((n_events + (1 << 20) - 1) >> 20) << 20;
    const int rcap = R <= 3 ? 3 : 8;
    auto key = std::make_tuple(device, H, W, cap, rcap);
    auto it = g_plans.find(key);
    if (it == g_plans.end()) {
        Entry e;
        const int rc = eincm_plan_create(&e.plan, device, H, W, cap, rcap, 0u);
        if (rc != EINCM_OK) { *err = eincm_last_error(nullptr); return nullptr; }
        it = g_plans.emplace(key, e).first;
    }
    return &it->second;
}

// Operands as loss_func receives them (losses.py:108-114): theta [h, w, 2] f64, xs / ys [N] i16, ts [N] f64, edges [R, H, W] f64,
// edge_ts [R] f64.  edge_ts become kernel constants, so the handler needs them on the host: they are passed as an attribute
// (`edge_ts` span) - the Python wrapper reads them from the NumPy array the loaders deliver.  `window_id` changes whenever
// set_datasample stages a new window (the wrapper bumps it), so the events are packed and sorted once per window, not per call.
ffi::Error ValueAndGradImpl(cudaStream_t stream, int32_t device, ffi::Buffer<ffi::F64> theta, ffi::Buffer<ffi::S16> xs,
                            ffi::Buffer<ffi::S16> ys, ffi::Buffer<ffi::F64> ts, ffi::Buffer<ffi::F64> edges,
                            ffi::Span<const double> edge_ts, double alpha, double beta, double gamma, double delta,
                            int32_t cur_pyr_lvl, int32_t n_pyr_lvls, int64_t window_id, ffi::ResultBuffer<ffi::F64> loss,
                            ffi::ResultBuffer<ffi::F64> grad) {
    const auto ed = edges.dimensions();       // [R, H, W]
    const auto td = theta.dimensions();       // [h, w, 2]
    if (ed.size() != 3 || td.size() != 3 || td[2] != 2) return ffi::Error::InvalidArgument("edges must be [R,H,W], theta [h,w,2]");
    const int R = (int)ed[0], H = (int)ed[1], W = (int)ed[2];
    if ((int64_t)edge_ts.size() != R) return ffi::Error::InvalidArgument("edge_ts attribute must hold R values");
    const int64_t n = (int64_t)xs.element_count();
    if ((int64_t)ys.element_count() != n || (int64_t)ts.element_count() != n)
        return ffi::Error::InvalidArgument("xs, ys, ts must have the same length");

    std::lock_guard<std::mutex> lock(g_mu);
    const char* err = nullptr;
    Entry* e = entry_for(device, H, W, n, R, &err);
    if (e == nullptr) return ffi::Error::Internal(err ? err : "eincm_plan_create failed");
    if (e->window_id != window_id) {
        // eincm_plan_set_window reads the staging totals back (one 12-byte copy): it synchronises `stream` once per window
        const int rc = eincm_plan_set_window(e->plan, xs.typed_data(), ys.typed_data(), ts.typed_data(), n, edges.typed_data(),
                                             edge_ts.begin(), R, stream);
        if (rc != EINCM_OK) return ffi::Error::Internal(eincm_last_error(e->plan));
        e->window_id = window_id;
    }
    eincm_hparams hp;
    std::memset(&hp, 0, sizeof(hp));
    hp.alpha = alpha; hp.beta = beta; hp.gamma = gamma; hp.delta = delta;
    hp.cur_pyr_lvl = cur_pyr_lvl; hp.n_pyr_lvls = n_pyr_lvls; hp.method = EINCM_METHOD_BILINEAR;
    const int rc = eincm_value_and_grad(e->plan, theta.typed_data(), (int)td[0], (int)td[1], &hp, loss->typed_data(),
                                        grad->typed_data(), stream);
    if (rc != EINCM_OK) return ffi::Error::Internal(eincm_last_error(e->plan));
    return ffi::Error::Success();             // asynchronous: XLA orders later work on `stream`
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(EincmValueAndGrad, ValueAndGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::DeviceOrdinal>()
                                  .Arg<ffi::Buffer<ffi::F64>>()    // theta
                                  .Arg<ffi::Buffer<ffi::S16>>()    // xs
                                  .Arg<ffi::Buffer<ffi::S16>>()    // ys
                                  .Arg<ffi::Buffer<ffi::F64>>()    // ts
                                  .Arg<ffi::Buffer<ffi::F64>>()    // edges
                                  .Attr<ffi::Span<const double>>("edge_ts")
                                  .Attr<double>("alpha")
                                  .Attr<double>("beta")
                                  .Attr<double>("gamma")
                                  .Attr<double>("delta")
                                  .Attr<int32_t>("cur_pyr_lvl")
                                  .Attr<int32_t>("n_pyr_lvls")
                                  .Attr<int64_t>("window_id")
                                  .Ret<ffi::Buffer<ffi::F64>>()    // loss []
                                  .Ret<ffi::Buffer<ffi::F64>>());  // grad [h, w, 2]
