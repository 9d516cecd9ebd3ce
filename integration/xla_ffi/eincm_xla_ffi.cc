// XLA FFI handlers that register the library's window staging and its objective + gradient as JAX custom calls on the CUDA platform.
//
// Two custom calls, because the identity of the staged window has to be a RUN-TIME quantity: jaxopt jits the objective once per
// pyramid level and re-uses the executable for every window (reference src/eincm/solver.py:165-173, :209-216), and every operand of
// the objective - `edge_ts` included - is a traced device buffer there.  Anything read at trace time (a Python global, an attribute)
// would be frozen into the cached executable.
//
//   eincm_set_window     (xs [N] i16, ys [N] i16, ts [N] f64, edges [R, H, W] f64, edge_ts [R] f64; attr slot) -> token [2] i64
//       issued EAGERLY from MultipleLevelEINCMSolver.set_datasample (solver.py:185-194), once per window, outside any jit.  Stages the
//       window into the plan of (device, slot) - eincm_plan_set_window_device_ts: events packed and sorted, zero-warp statistics; the
//       reference times are copied device -> host here, once per window - and returns {slot, generation} as a device array.
//   eincm_value_and_grad (theta [h, w, 2] f64, token [2] i64; attrs hyper-parameters, slot) -> loss [] f64, grad [h, w, 2] f64
//       the body of jit(value_and_grad(loss_func)) (reference src/eincm/losses.py:108-205).  The token operand is the data dependency
//       that ties an evaluation to the staging call it follows; the window itself is whatever eincm_set_window staged last on (device,
//       slot) - both calls are enqueued on XLA's compute stream in program order.  Only enqueues: XLA owns buffers and stream.
//
// `slot` distinguishes solvers that share a device (one plan each); it is a static attribute - a solver instance owns its jitted
// executables, so a trace-time constant is exactly right for it.
//
// BUILD: needs the XLA FFI headers that ship with jaxlib (`python -c "import jax; print(jax.ffi.include_dir())"`); jax / jaxlib are not
// installed in this repository's image (no network), so the real handler has not been loaded by XLA here.  The SAME source is compiled
// against a minimal stand-in for `xla/ffi/api/ffi.h` (tests/native/mock_xla_ffi) and its handler bodies are driven on a GPU by
// tests/test_gpu_ffi_handler.py: two consecutive windows of identical shape through SetWindowImpl / ValueAndGradImpl, results compared
// with the C-ABI called directly.  Where jaxlib exists:
//
//   g++ -O2 -std=c++17 -shared -fPIC eincm_xla_ffi.cc -o libeincm_xla_ffi.so
//       -I"$(python -c 'import jax; print(jax.ffi.include_dir())')" -I../../include -I/usr/local/cuda/include
//       -L../../edge-informed-contrast-maximization_b200/lib -leincm_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN'
//
// Python side: integration/xla_ffi/eincm_jax.py.
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>

#include <cuda_runtime_api.h>

#include "eincm.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

// One plan per (device, slot).  XLA may call a handler from any host thread: the table and every plan are guarded by one mutex
// (plans are thread-compatible, not thread-safe; an evaluation is a handful of launches, the critical section is microseconds).
struct Entry {
    eincm_plan* plan = nullptr;
    int H = 0, W = 0, max_refs = 0;
    int64_t max_events = 0;
    int64_t generation = 0;             // windows staged so far; 0: none
    int64_t* token_host = nullptr;      // pinned: source of the asynchronous copy into the token result
};
std::mutex g_mu;
std::map<std::pair<int, int>, Entry> g_entries;

ffi::Error fail(const eincm_plan* plan, const char* what) {
    const char* m = eincm_last_error(plan);
    return ffi::Error::Internal(std::string(what) + ": " + (m ? m : "unknown error"));
}

ffi::Error SetWindowImpl(cudaStream_t stream, int32_t device, ffi::Buffer<ffi::S16> xs, ffi::Buffer<ffi::S16> ys, ffi::Buffer<ffi::F64> ts,
                         ffi::Buffer<ffi::F64> edges, ffi::Buffer<ffi::F64> edge_ts, int32_t slot, ffi::ResultBuffer<ffi::S64> token) {
    const auto ed = edges.dimensions();       // [R, H, W]
    if (ed.size() != 3) return ffi::Error::InvalidArgument("edges must be [R, H, W]");
    const int R = (int)ed[0], H = (int)ed[1], W = (int)ed[2];
    if ((int64_t)edge_ts.element_count() != R) return ffi::Error::InvalidArgument("edge_ts must hold one reference time per edge image");
    const int64_t n = (int64_t)xs.element_count();
    if ((int64_t)ys.element_count() != n || (int64_t)ts.element_count() != n)
        return ffi::Error::InvalidArgument("xs, ys, ts must have the same length");
    if (token->element_count() != 2) return ffi::Error::InvalidArgument("the token result must be int64[2]");

    std::lock_guard<std::mutex> lock(g_mu);
    Entry& e = g_entries[std::make_pair((int)device, (int)slot)];
    if (e.plan == nullptr || e.H != H || e.W != W || n > e.max_events || R > e.max_refs) {
        // first window of this slot, or one that does not fit: capacities are rounded up so that the windows of a sequence share a plan
        if (e.plan != nullptr) { eincm_plan_destroy(e.plan); e.plan = nullptr; }
        const int64_t cap = ((n + (1 << 20) - 1) >> 20) << 20;
        const int rcap = R <= 3 ? 3 : 8;                       // the library holds up to 8 reference times per window
        if (eincm_plan_create(&e.plan, device, H, W, cap > 0 ? cap : (1 << 20), rcap, 0u) != EINCM_OK) { e.plan = nullptr; return fail(nullptr, "eincm_plan_create"); }
        e.H = H; e.W = W; e.max_events = cap > 0 ? cap : (1 << 20); e.max_refs = rcap;
    }
    if (e.token_host == nullptr && cudaMallocHost((void**)&e.token_host, 2 * sizeof(int64_t)) != cudaSuccess)
        return ffi::Error::Internal("cudaMallocHost failed");
    // synchronises `stream` (reference times device -> host, staging totals read back): once per window, in an eager call - never
    // inside the jitted objective
    if (eincm_plan_set_window_device_ts(e.plan, xs.typed_data(), ys.typed_data(), ts.typed_data(), n, edges.typed_data(), edge_ts.typed_data(), R,
                                        stream) != EINCM_OK)
        return fail(e.plan, "eincm_plan_set_window");
    ++e.generation;
    e.token_host[0] = slot; e.token_host[1] = e.generation;
    if (cudaMemcpyAsync(token->typed_data(), e.token_host, 2 * sizeof(int64_t), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess)            // token_host is rewritten by the next window
        return ffi::Error::Internal("copy of the window token failed");
    return ffi::Error::Success();
}

ffi::Error ValueAndGradImpl(cudaStream_t stream, int32_t device, ffi::Buffer<ffi::F64> theta, ffi::Buffer<ffi::S64> token, double alpha,
                            double beta, double gamma, double delta, int32_t cur_pyr_lvl, int32_t n_pyr_lvls, int32_t slot,
                            ffi::ResultBuffer<ffi::F64> loss, ffi::ResultBuffer<ffi::F64> grad) {
    const auto td = theta.dimensions();       // [h, w, 2]
    if (td.size() != 3 || td[2] != 2) return ffi::Error::InvalidArgument("theta must be [h, w, 2]");
    if (token.element_count() != 2) return ffi::Error::InvalidArgument("the window token must be int64[2] (the result of eincm_set_window)");
    if (grad->element_count() != theta.element_count()) return ffi::Error::InvalidArgument("grad must have the shape of theta");

    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_entries.find(std::make_pair((int)device, (int)slot));
    if (it == g_entries.end() || it->second.plan == nullptr || it->second.generation == 0)
        return ffi::Error::Internal("eincm_value_and_grad before eincm_set_window on this device / slot");
    eincm_hparams hp;
    std::memset(&hp, 0, sizeof(hp));
    hp.alpha = alpha; hp.beta = beta; hp.gamma = gamma; hp.delta = delta;
    hp.cur_pyr_lvl = cur_pyr_lvl; hp.n_pyr_lvls = n_pyr_lvls; hp.method = EINCM_METHOD_BILINEAR;
    if (eincm_value_and_grad(it->second.plan, theta.typed_data(), (int)td[0], (int)td[1], &hp, loss->typed_data(), grad->typed_data(), stream) != EINCM_OK)
        return fail(it->second.plan, "eincm_value_and_grad");
    return ffi::Error::Success();             // asynchronous: XLA orders later work on `stream`
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(EincmSetWindow, SetWindowImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::DeviceOrdinal>()
                                  .Arg<ffi::Buffer<ffi::S16>>()    // xs
                                  .Arg<ffi::Buffer<ffi::S16>>()    // ys
                                  .Arg<ffi::Buffer<ffi::F64>>()    // ts
                                  .Arg<ffi::Buffer<ffi::F64>>()    // edges [R, H, W]
                                  .Arg<ffi::Buffer<ffi::F64>>()    // edge_ts [R]: a device buffer like every other operand
                                  .Attr<int32_t>("slot")
                                  .Ret<ffi::Buffer<ffi::S64>>());  // token [2]

XLA_FFI_DEFINE_HANDLER_SYMBOL(EincmValueAndGrad, ValueAndGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::DeviceOrdinal>()
                                  .Arg<ffi::Buffer<ffi::F64>>()    // theta [h, w, 2]
                                  .Arg<ffi::Buffer<ffi::S64>>()    // token [2]
                                  .Attr<double>("alpha")
                                  .Attr<double>("beta")
                                  .Attr<double>("gamma")
                                  .Attr<double>("delta")
                                  .Attr<int32_t>("cur_pyr_lvl")
                                  .Attr<int32_t>("n_pyr_lvls")
                                  .Attr<int32_t>("slot")
                                  .Ret<ffi::Buffer<ffi::F64>>()    // loss []
                                  .Ret<ffi::Buffer<ffi::F64>>());  // grad [h, w, 2]
