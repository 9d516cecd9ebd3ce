// C-ABI of include/eincm.h: the plan object and the launch sequences of one window / one evaluation.
// Host side only orchestrates; all arithmetic is in the k_*.cuh kernels (sm_100a).  No CPU fallback exists:
// every compute entry point fails with EINCM_ECUDA when no CUDA device of compute capability 10.x is usable.
#include "../../include/eincm.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <map>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>
#include <vector>

#include "common.cuh"
#include "k_edges.cuh"
#include "k_preproc.cuh"
#include "k_ingest.cuh"
#include "k_prep.cuh"
#include "k_theta.cuh"
#include "k_opt.cuh"
#include "k_events.cuh"
#include "k_events_tile.cuh"
#include "k_image.cuh"
#include "k_eval.cuh"
#include "k_image_fused.cuh"
#include "eincm_opt.h"

using namespace eincm;

namespace {

thread_local std::string g_create_error;

struct AxisTapsOwner {
    int n_in = 0, n_out = 0;
    void* blob = nullptr;   // one device allocation holding all six arrays
    AxisTaps taps{};
};

}  // namespace

struct eincm_plan {
    int device = 0, H = 0, W = 0, max_refs = 0, sm_count = 148;
    int64_t HW = 0, max_events = 0;
    unsigned flags = 0;
    bool wrap = true, exact = false;
    int split_rank = 0, split_world = 1;   // event-split plans: rank 0 alone adds the (replicated) TV gradient
    // window state
    int64_t n_events = 0;
    int64_t n_stream = 0, stream_cap = 0;   // padded length of the sorted stream (k_prep.cuh) / its capacity
    int n_chunks = 0, chunk_cap = 0, n_tiles = 0;
    int R = 0;
    bool window_set = false, window_final = false, zero_div_valid = false, forward_done = false;
    ThetaSrc tsrc{};              // flow operand of the last forward pass
    bool theta_full_valid = false;
    unsigned long long* peer_fix[kMaxPeers] = {};   // event split with peer access: fixed-point images of all ranks (own included)
    void* peer_opened[kMaxPeers] = {};              // pointers obtained from cudaIpcOpenMemHandle (closed on destroy)
    int n_peers = 0;                                // 0: no peer access (the caller all-reduces the images)
    bool split_fixed = false;                       // event split without peer access: the caller all-reduces the FIXED-POINT images (int64)
    eincm_group* group = nullptr;        // evaluation group this plan rendezvous with inside eincm_minimize_bfgs_host (or null)
    cudaEvent_t wait_ev = nullptr;       // EINCM_FLAG_BLOCKING_SYNC: blocking-sync event the host entry points sleep on
    cudaStream_t own_stream = nullptr;   // for the batched host call: every plan of a batch runs on its own stream
    double host_enqueue_s = 0.0, host_wait_s = 0.0;   // wall time spent launching / waiting in the synchronous host entry points
    int64_t host_evals = 0;
    double host_seq = 0.0;               // sequence number of the last host evaluation (written back by its last kernel)
    size_t host_ng = 0;                  // gradient doubles of the host evaluation in flight (host_enqueue -> host_collect)
    bool host_delivered = false;  // the last backward pass wrote its results into the mapped host buffer itself
    double* h_mapped_dev = nullptr;   // device alias of h_pinned (cudaHostAllocMapped)
    bool dldi_stale = false;      // plan->dldi does not hold the last evaluation's d loss / d IWE yet (fused path: built on demand)
    bool fix_clean = false;       // every cell of iwe_fix is zero (the cooperative image pass clears what it reads)
    bool coop_ok = false;         // the sensor is narrow enough for the row-band cooperative image pass
    bool fused_pending = false;   // the last forward left the fixed-point images for the fused image pass (no float64 copy yet)
    RefTimes tref{};
    // last evaluation
    int last_h = 0, last_w = 0;
    const double* last_theta = nullptr;   // device pointers of the theta operands of the last forward
    const double* last_prev = nullptr;
    double last_a_ho = 0.0;
    // device buffers
    uint32_t* ev_xy = nullptr;
    double* ev_t = nullptr;
    uint32_t* perm = nullptr;
    double* ev_t2 = nullptr;      // ping-pong partners of ev_t / perm for the per-pixel time ordering
    uint32_t* perm2 = nullptr;
    unsigned int *counts = nullptr, *cursor = nullptr, *tile_cnt = nullptr, *tile_start = nullptr, *chunk_first = nullptr, *totals = nullptr;
    Chunk* chunks = nullptr;
    Chunk* chunks2 = nullptr;                          // second table: k_chunk_order writes the size-ordered table here, then the two swap
    int4* chunk_win = nullptr;                         // [chunk_cap][max_refs] windows of the last forward pass
    float* adj32 = nullptr;                            // [max_refs][H*W] adjoint of the Scharr pair applied to (Gx, Gy) (k_image_stats -> k_image_grad)
    float2* chunk_tr = nullptr;                        // [chunk_cap] t range of the events of every chunk (per window)
    unsigned long long* iwe_fix = nullptr;             // [max_refs][H*W] fixed-point images of warped events
    int n_keys = 0, tiles_x = 0;
    uint8_t* mask = nullptr;
    double2 *theta_full = nullptr, *Gtv = nullptr, *partial = nullptr;
    double *G = nullptr, *iwe = nullptr, *zero_iwe = nullptr, *dldi = nullptr, *edges = nullptr;
    float* dldi32 = nullptr;                           // default path: float32 copy of dL/dIWE, scaled by 1/(2 pi)
    cudaEvent_t window_ev = nullptr;                   // recorded behind the last kernel of set_window / window_finalize
    cudaStream_t window_stream = nullptr;              // ... on this stream; evaluations on another stream wait for it once
    bool window_ev_valid = false;
    cudaStream_t window_waited = nullptr;              // stream that has already waited for window_ev of the current window
    bool window_waited_valid = false;
    double *sbar = nullptr, *gNdiv = nullptr;          // delta != 0 only, allocated on first use
    double* part = nullptr;                            // per-CTA partials of the two-level reductions
    int part_doubles = 0;
    DevScalars* sc = nullptr;
    double *theta_stage = nullptr, *prev_stage = nullptr, *grad_stage = nullptr, *grad_buf = nullptr, *out_stage = nullptr;
    int16_t *xs_stage = nullptr, *ys_stage = nullptr;  // stateless host form only, allocated on first use
    double* ts_stage = nullptr;
    double* edges_stage = nullptr;
    // pinned host staging
    double* h_pinned = nullptr;     // [2*H*W + 1024]
    int* h_flag = nullptr;
    std::map<std::pair<int, int>, AxisTapsOwner> taps_cache;
    // device-side solve loop (eincm_minimize_bfgs_graph_host, k_opt.cuh): state, vectors, dense inverse Hessian, one graph per level key
    BfgsDev* bfgs_state = nullptr;
    double* bfgs_vec = nullptr;          // x, g, p, s, y, Hy, x_trial, g_trial [n each], f_trial, result [4 + n]
    double* bfgs_H = nullptr;
    int bfgs_cap_n = 0;
    uint64_t window_gen = 0;             // staged windows so far: graphs bake per-window kernel parameters (reference times, chunk count)
    struct LevelGraph { cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr; uint64_t gen = 0; };
    std::map<std::vector<double>, LevelGraph> bfgs_graphs;
    const int* eval_skip = nullptr;      // while a solve graph is captured: the level's "done" flag, handed to the evaluation kernels
    int graph_unroll = 8;                // evaluations per launch of an unrolled solve graph; 0: the WHILE conditional form (EINCM_GRAPH_UNROLL)
    bool graph_no_pdl = false;           // programmatic dependent launches could not be captured into the loop body: plain launches there
    std::string error;
    // launch accounting / optional per-kernel timing
    int64_t launch_count = 0;
    bool timing = false;
    struct Span { const char* name; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
};

namespace {

int fail(eincm_plan* p, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (p) p->error = buf; else g_create_error = buf;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(plan, e_ == cudaErrorMemoryAllocation ? EINCM_ENOMEM : EINCM_ECUDA, "%s: %s",   \
                        #call, cudaGetErrorString(e_));                                                  \
    } while (0)

// One kernel launch on stream `st` of `plan`: counts it, checks the launch, and - when timing is enabled
// (eincm_plan_set_timing) - brackets it with CUDA events on the same stream.
#define LAUNCH(name, ...)                                                                                \
    do {                                                                                                 \
        const int span_ = span_begin(plan, name, st);                                                    \
        __VA_ARGS__;                                                                                     \
        cudaError_t e_ = cudaGetLastError();                                                             \
        if (e_ != cudaSuccess) return fail(plan, EINCM_ECUDA, "launch %s: %s", name, cudaGetErrorString(e_)); \
        span_end(plan, span_, st);                                                                       \
    } while (0)

// Launch with programmatic stream serialization (PDL): the kernel may be scheduled while the previous kernel of the stream drains;
// it calls griddepcontrol.wait before it touches anything the previous kernel produces.
thread_local bool t_plain_launches = false;      // set while a loop body is captured without programmatic edges

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = t_plain_launches ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

int span_begin(eincm_plan* p, const char* name, cudaStream_t st) {
    ++p->launch_count;
    if (!p->timing) return -1;
    cudaEvent_t ev[2];
    for (auto& e : ev) {
        if (!p->event_pool.empty()) { e = p->event_pool.back(); p->event_pool.pop_back(); }
        else if (cudaEventCreate(&e) != cudaSuccess) return -1;
    }
    cudaEventRecord(ev[0], st);
    p->spans.push_back({name, ev[0], ev[1]});
    return (int)p->spans.size() - 1;
}

void span_end(eincm_plan* p, int idx, cudaStream_t st) {
    if (idx >= 0) cudaEventRecord(p->spans[idx].b, st);
}

template <typename T>
cudaError_t dmalloc(T** p, size_t count) { return cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)); }

// Per-axis resize taps of jax.image.scale_and_translate(method='bilinear', antialias=True) for n_in -> n_out
// (n_out >= n_in), restating compute_weight_mat (reference src/utils/theta_utils.py:25-35).
int build_axis_taps(eincm_plan* plan, int n_in, int n_out, AxisTaps* out) {
    auto key = std::make_pair(n_in, n_out);
    auto it = plan->taps_cache.find(key);
    if (it != plan->taps_cache.end()) { *out = it->second.taps; return EINCM_OK; }
    std::vector<int> i0(n_out), i1(n_out), lo(n_in, n_out), hi(n_in, 0);
    std::vector<double> w0(n_out), w1(n_out);
    const double scale = (double)n_out / (double)n_in;
    const double inv = 1.0 / scale;
    const double kscale = std::max(inv, 1.0);
    for (int j = 0; j < n_out; ++j) {
        const double f = ((double)j + 0.5) * inv - 0.0 * inv - 0.5;
        int a = (int)std::floor(f), b = a + 1;
        double wa = 0.0, wb = 0.0;
        if (a >= 0 && a < n_in) wa = std::max(0.0, 1.0 - std::fabs(f - (double)a) / kscale);
        if (b >= 0 && b < n_in) wb = std::max(0.0, 1.0 - std::fabs(f - (double)b) / kscale);
        const double total = wa + wb;
        if (std::fabs(total) > 1000.0 * 1.1920928955078125e-07) { wa /= total; wb /= total; } else { wa = wb = 0.0; }
        if (!(f >= -0.5 && f <= (double)n_in - 0.5)) wa = wb = 0.0;
        if (a < 0 || a >= n_in) { a = b; wa = wb; wb = 0.0; }          // single valid tap: keep it in slot 0
        if (b < 0 || b >= n_in || wb == 0.0) { b = a; wb = 0.0; }
        if (a < 0 || a >= n_in) { a = b = 0; wa = wb = 0.0; }
        i0[j] = a; i1[j] = b; w0[j] = wa; w1[j] = wb;
        if (wa != 0.0) { lo[a] = std::min(lo[a], j); hi[a] = std::max(hi[a], j + 1); }
        if (wb != 0.0) { lo[b] = std::min(lo[b], j); hi[b] = std::max(hi[b], j + 1); }
    }
    for (int i = 0; i < n_in; ++i) if (hi[i] <= lo[i]) { lo[i] = 0; hi[i] = 0; }
    AxisTapsOwner o;
    o.n_in = n_in; o.n_out = n_out;
    const size_t bytes = (size_t)n_out * (2 * sizeof(int) + 2 * sizeof(double)) + (size_t)n_in * 2 * sizeof(int);
    CU(cudaMalloc(&o.blob, bytes));
    char* base = (char*)o.blob;
    double* d_w0 = (double*)base; double* d_w1 = d_w0 + n_out;
    int* d_i0 = (int*)(d_w1 + n_out); int* d_i1 = d_i0 + n_out; int* d_lo = d_i1 + n_out; int* d_hi = d_lo + n_in;
    // synchronous copies from pageable memory: happens once per (n_in, n_out) pair and plan
    CU(cudaMemcpy(d_w0, w0.data(), n_out * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_w1, w1.data(), n_out * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_i0, i0.data(), n_out * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_i1, i1.data(), n_out * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_lo, lo.data(), n_in * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_hi, hi.data(), n_in * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaDeviceSynchronize());   // pageable H2D may still be in flight w.r.t. non-blocking streams
    o.taps = AxisTaps{d_i0, d_i1, d_w0, d_w1, d_lo, d_hi};
    plan->taps_cache[key] = o;
    *out = o.taps;
    return EINCM_OK;
}

__global__ void k_set_weights(DevScalars* sc, RefTimes w, int R) {
    if (threadIdx.x < EINCM_MAX_REFS) sc->weights[threadIdx.x] = threadIdx.x < R ? w.t[threadIdx.x] : 0.0;
}

int event_grid(const eincm_plan* p, int64_t n, int threads) {
    const int64_t want = (n + threads - 1) / threads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)p->sm_count * 8));
}

int img_tiles_x(const eincm_plan* p) { return (p->W + kImgTX - 1) / kImgTX; }
int img_tiles_y(const eincm_plan* p) { return (p->H + kImgTY - 1) / kImgTY; }
int img_nb_flat(const eincm_plan* p) { return (int)std::min<int64_t>((p->HW + kImgNT - 1) / kImgNT, (int64_t)p->sm_count * 4); }

int check_hp(eincm_plan* plan, const eincm_hparams* hp) {
    if (!hp) return fail(plan, EINCM_EINVAL, "hparams is NULL");
    if (hp->method != EINCM_METHOD_BILINEAR)
        return fail(plan, EINCM_EUNSUPPORTED, "scale_to_sensor_size_method %d unsupported (only bilinear, the shipped default)", hp->method);
    return EINCM_OK;
}

int ensure_delta_buffers(eincm_plan* plan) {
    if (plan->sbar) return EINCM_OK;
    CU(dmalloc(&plan->sbar, (size_t)plan->max_refs * plan->HW));
    CU(dmalloc(&plan->gNdiv, (size_t)plan->max_refs * plan->HW));
    return EINCM_OK;
}

// zero-IWE statistics that depend on delta (computed on first use; losses.py:80)
int ensure_zero_div(eincm_plan* plan, cudaStream_t st) {
    if (plan->zero_div_valid) return EINCM_OK;
    int rc = ensure_delta_buffers(plan);
    if (rc) return rc;
    const dim3 grid(img_tiles_x(plan), img_tiles_y(plan), 1), block(kImgTX, kImgTY);
    LAUNCH("k_img_D1(zero)", k_img_D1<<<grid, block, 0, st>>>(plan->zero_iwe, 0, plan->H, plan->W, grid.x * grid.y, plan->sc->zero, 0, nullptr, nullptr,
                                     plan->part, plan->sc->zero, &plan->sc->counters[2]));
    plan->zero_div_valid = true;
    return EINCM_OK;
}

int window_finalize_impl(eincm_plan* plan, cudaStream_t st) {
    const dim3 block(kImgTX, kImgTY);
    const dim3 gridA(img_tiles_x(plan), img_tiles_y(plan), 1);
    LAUNCH("k_img_A(zero)", k_img_A<<<gridA, block, 0, st>>>(plan->zero_iwe, plan->H, plan->W, gridA.x * gridA.y, plan->part, plan->sc->zero,
                                     &plan->sc->counters[0]));
    const int nb = img_nb_flat(plan);
    LAUNCH("k_img_B(zero)", k_img_B<<<dim3(nb, 1, plan->R), block, 0, st>>>(plan->zero_iwe, 0, plan->edges, nullptr, plan->HW, nb, plan->sc->zero, 0,
                                                    nullptr, plan->part, plan->sc->zero, &plan->sc->counters[1]));
    plan->window_final = true;
    plan->zero_div_valid = false;
    // split-phase use (eincm_window_finalize is asynchronous): evaluations enqueued on ANOTHER stream must not overtake these
    // kernels - they wait for this event once (forward_events_impl)
    if (cudaEventRecord(plan->window_ev, st) != cudaSuccess) return fail(plan, EINCM_ECUDA, "cudaEventRecord failed");
    plan->window_ev_valid = true; plan->window_stream = st; plan->window_waited_valid = false;
    return EINCM_OK;
}

// Builds n_img images of warped events (one per reference time in `tref`) from the staged events: exact mode = nine
// float64 scatter-adds per event and image; default = tile-privatised fixed-point splat (k_events_tile.cuh), optionally
// converted to a float64 image.
int splat_images(eincm_plan* plan, const ThetaSrc& T, const double2* theta_full, int n_img, const RefTimes& tref, double* out, const char* tag,
                 cudaStream_t st, bool to_f64 = true, bool record_windows = false) {
    const int64_t n = plan->n_events > 0 ? plan->n_stream : 0;      // padded stream: sentinels are skipped by the kernels
    const int H = plan->H, W = plan->W;
    if (plan->exact) {
        CU(cudaMemsetAsync(out, 0, (size_t)n_img * plan->HW * sizeof(double), st));
        if (n > 0) {
            const int grid = event_grid(plan, n, 256);
            if (plan->wrap) LAUNCH(tag, k_splat<true><<<grid, 256, 0, st>>>(plan->ev_xy, plan->ev_t, n, theta_full, H, W, n_img, tref, out));
            else LAUNCH(tag, k_splat<false><<<grid, 256, 0, st>>>(plan->ev_xy, plan->ev_t, n, theta_full, H, W, n_img, tref, out));
        }
        return EINCM_OK;
    }
    if (plan->n_peers > 0) {
        // peers add into this rank's image: it must be clean BEFORE the cross-rank barrier that precedes the splat
        if (!plan->fix_clean) return fail(plan, EINCM_ESTATE, "event split with peer access: call eincm_split_prepare (+ barrier) before the splat");
    } else if (!plan->fix_clean) {
        CU(cudaMemsetAsync(plan->iwe_fix, 0, (size_t)plan->max_refs * plan->HW * sizeof(unsigned long long), st));
    }
    plan->fix_clean = false;
    if (n > 0) {
        const int grid = std::max(1, plan->n_chunks);
        int4* cw = record_windows ? plan->chunk_win : nullptr;
        FixDst dst{};
        if (plan->n_peers > 0) { dst.n = plan->n_peers; for (int q = 0; q < dst.n; ++q) dst.p[q] = plan->peer_fix[q]; }
        else { dst.n = 1; dst.p[0] = plan->iwe_fix; }
#define SPLATT(WR, RB) LAUNCH(tag, launch_pdl(k_splat_tile<WR, RB>, dim3(grid), dim3(256), RB * kWinCap * sizeof(uint32_t), st, (const uint32_t*)plan->ev_xy, \
                               (const double*)plan->ev_t, (const Chunk*)plan->chunks, (const float2*)plan->chunk_tr, (const unsigned int*)(plan->totals + 1), T, H, W, n_img, tref, dst, cw, plan->eval_skip))
#define SPLATT_RB(WR) do { switch (refs_per_pass(n_img)) { case 1: SPLATT(WR, 1); break; case 2: SPLATT(WR, 2); break; \
                                                             case 3: SPLATT(WR, 3); break; default: SPLATT(WR, 4); } } while (0)
        if (plan->wrap) SPLATT_RB(true); else SPLATT_RB(false);
#undef SPLATT_RB
#undef SPLATT
    }
    if (!to_f64) return EINCM_OK;
    const int64_t cells = (int64_t)n_img * plan->HW;
    LAUNCH("k_fix_to_f64", k_fix_to_f64<<<(int)std::min<int64_t>((cells + 255) / 256, plan->sm_count * 8), 256, 0, st>>>(plan->iwe_fix, cells, out));
    return EINCM_OK;
}

// dense flow field of the last forward pass (k_upsample_theta), computed at most once per evaluation
int ensure_theta_full(eincm_plan* plan, cudaStream_t st) {
    if (plan->theta_full_valid) return EINCM_OK;
    const dim3 block(32, 8), grid((plan->W + 31) / 32, (plan->H + 7) / 8);
    LAUNCH("k_upsample_theta", k_upsample_theta<<<grid, block, 0, st>>>(plan->tsrc, plan->H, plan->W, plan->theta_full));
    plan->theta_full_valid = true;
    return EINCM_OK;
}

int forward_events_impl(eincm_plan* plan, const double* theta, const double* prev, double a_ho, int h, int w,
                        const eincm_hparams* hp, cudaStream_t st) {
    int rc = check_hp(plan, hp);
    if (rc) return rc;
    if (!plan->window_set) return fail(plan, EINCM_ESTATE, "value_and_grad before set_window");
    if (!theta) return fail(plan, EINCM_EINVAL, "theta is NULL");
    if (h < 1 || w < 1 || h > plan->H || w > plan->W)
        return fail(plan, EINCM_EINVAL, "theta shape (%d,%d) must be within (1,1)..(%d,%d): the resize only up-scales", h, w, plan->H, plan->W);
    AxisTaps ty, tx;
    if ((rc = build_axis_taps(plan, h, plan->H, &ty))) return rc;
    if ((rc = build_axis_taps(plan, w, plan->W, &tx))) return rc;
    if (plan->window_ev_valid && st != plan->window_stream && !(plan->window_waited_valid && plan->window_waited == st)) {
        CU(cudaStreamWaitEvent(st, plan->window_ev, 0));
        plan->window_waited = st; plan->window_waited_valid = true;
    }
    plan->tsrc = ThetaSrc{theta, prev, a_ho, h, w, ty, tx};
    plan->theta_full_valid = false;
    // the dense field is only materialised for the float64 nine-tap kernels and the TV regulariser (and, on demand, the debug tap);
    // the default event kernels evaluate theta per source tile
    if (plan->exact || (hp->gamma != 0.0 && hp->cur_pyr_lvl <= 0)) {
        if ((rc = ensure_theta_full(plan, st))) return rc;
    }
    // default path (single GPU, delta == 0): the fixed-point images are consumed by the fused image pass directly
    // single GPU, or event split with peer access (every rank then holds the complete fixed-point images after the barrier)
    plan->fused_pending = !plan->exact && plan->coop_ok && (!(plan->flags & EINCM_FLAG_EVENT_SPLIT) || plan->n_peers > 0 || plan->split_fixed) && hp->delta == 0.0;
    if ((rc = splat_images(plan, plan->tsrc, plan->theta_full, plan->R, plan->tref, plan->iwe, "k_splat", st, !plan->fused_pending, true))) return rc;
    plan->last_h = h; plan->last_w = w; plan->last_theta = theta; plan->last_prev = prev; plan->last_a_ho = a_ho;
    plan->forward_done = true;
    return EINCM_OK;
}

// host_out: device alias of mapped pinned memory; when the tile-theta gradient kernel runs it delivers [grad | loss | dalpha]
// there itself (plan->host_delivered), otherwise the caller copies the results back.
int backward_impl(eincm_plan* plan, const eincm_hparams* hp, double* loss_out, double* grad_out, double* dalpha_out, cudaStream_t st,
                  double* host_out = nullptr) {
    plan->host_delivered = false;
    plan->dldi_stale = false;
    int rc = check_hp(plan, hp);
    if (rc) return rc;
    if (!plan->forward_done) return fail(plan, EINCM_ESTATE, "eincm_backward before eincm_forward_events");
    if (!plan->window_final) return fail(plan, EINCM_ESTATE, "eincm_backward before eincm_window_finalize");
    const int R = plan->R, H = plan->H, W = plan->W, h = plan->last_h, w = plan->last_w;
    const bool use_tv = hp->gamma != 0.0 && hp->cur_pyr_lvl <= 0;       // losses.py:171
    const bool use_div = hp->delta != 0.0;
    const bool want_grad = grad_out != nullptr || dalpha_out != nullptr;
    bool g_zeroed = false;               // the cooperative image pass also clears G
    bool g_zeroed_grad = false;          // ... and the theta-gradient accumulator of k_theta_grad
    const dim3 block(kImgTX, kImgTY);
    const dim3 gridT(img_tiles_x(plan), img_tiles_y(plan), R);
    const int nbT = gridT.x * gridT.y;
    if (plan->fused_pending && use_div) {        // hparams changed between the split-phase calls: fall back to the unfused pass
        const int64_t cells = (int64_t)R * plan->HW;
        LAUNCH("k_fix_to_f64", k_fix_to_f64<<<(int)std::min<int64_t>((cells + 255) / 256, plan->sm_count * 8), 256, 0, st>>>(plan->iwe_fix, cells, plan->iwe));
        plan->fused_pending = false;
    }
    if (plan->fused_pending) {
        if (use_tv) {
            const dim3 gridV((W + kTvTX - 1) / kTvTX, (H + kTvTY - 1) / kTvTY), blockV(kTvTX, kTvTY);
            if ((rc = ensure_theta_full(plan, st))) return rc;
            LAUNCH("k_tv", k_tv<<<gridV, blockV, 0, st>>>(plan->theta_full, plan->mask, H, W, gridV.x * gridV.y, plan->Gtv, plan->part, plan->sc));
        }
        const int tvb = ((W + kTvTX - 1) / kTvTX) * ((H + kTvTY - 1) / kTvTY);
        ImageStatsArgs ia{};
        ia.fix = plan->iwe_fix; ia.edges = plan->edges; ia.iwe = plan->iwe; ia.adj32 = plan->adj32;
        ia.part = plan->part + 2 * tvb;                          // k_tv's partials live at the start of `part`
        ia.sc = plan->sc; ia.loss_out = loss_out;
        ia.skip = plan->eval_skip;
        ia.zero_buf = want_grad ? plan->G : nullptr; ia.n_zero = (int)(plan->HW * 2);     // cleared for the event backward pass
        g_zeroed = want_grad;
        if (want_grad && h * w <= kGatherMaxTiles) {
            ia.zero_buf2 = grad_out ? grad_out : plan->grad_buf; ia.n_zero2 = h * w * 2;
            g_zeroed_grad = true;
        }
        ia.H = H; ia.W = W; ia.R = R;
        ia.alpha = hp->alpha; ia.beta = hp->beta; ia.gamma = hp->gamma; ia.use_tv = use_tv ? 1 : 0;
        LAUNCH("k_image_stats", launch_pdl(k_image_stats, dim3(R * image_stats_ctas(H, W)), dim3(kS2NT), 0, st, ia));
        // the loss is evaluated by a spare CTA of k_image_grad (publish_loss), off the critical path of the evaluation
        auto tail_args = [&](ImageGradArgs& ga) {
            ga.publish = 1; ga.skip = plan->eval_skip; ga.loss_out = loss_out; ga.alpha = hp->alpha; ga.beta = hp->beta; ga.gamma = hp->gamma; ga.use_tv = use_tv ? 1 : 0;
        };
        ImageGradArgs ga{};
        ga.fix = plan->iwe_fix; ga.edges = plan->edges; ga.iwe = plan->iwe; ga.adj32 = plan->adj32; ga.sc = plan->sc;
        ga.dldi = nullptr; ga.dldi32 = plan->dldi32; ga.HW = (int)plan->HW; ga.R = R; ga.want_grad = want_grad ? 1 : 0;
        tail_args(ga);
        plan->dldi_stale = want_grad;                            // the float64 copy (debug tap) is produced on demand: eincm_dldi_ptr
        LAUNCH("k_image_grad", launch_pdl(k_image_grad, dim3(std::max(1, std::min((int)((plan->HW + 1023) / 1024), plan->sm_count * 8)) + 1), dim3(256), 0, st, ga));   // + 1: the publishing CTA
        plan->fused_pending = false;
        plan->fix_clean = true;                                  // the pass clears the cells it has read
        if (!want_grad) return EINCM_OK;
    } else {
        if (use_div) {
            if ((rc = ensure_zero_div(plan, st))) return rc;
        }
        LAUNCH("k_img_A", k_img_A<<<gridT, block, 0, st>>>(plan->iwe, H, W, nbT, plan->part, plan->sc->ref, &plan->sc->counters[0]));
        LAUNCH("k_scalars(0)", k_scalars<<<1, 32, 0, st>>>(plan->sc, R, (double)plan->HW, hp->alpha, hp->beta, hp->gamma, hp->delta, use_tv, use_div, 0, nullptr));
        const double* gNdiv = nullptr;
        if (use_div) {
            LAUNCH("k_img_D1", k_img_D1<<<gridT, block, 0, st>>>(plan->iwe, plan->HW, H, W, nbT, plan->sc->ref, 1, plan->sc->coefD, plan->sbar, plan->part,
                                              plan->sc->ref, &plan->sc->counters[2]));
            if (want_grad) {
                LAUNCH("k_img_D2", k_img_D2<<<gridT, block, 0, st>>>(plan->sbar, H, W, plan->gNdiv));
                gNdiv = plan->gNdiv;
            }
        }
        const int nbF = img_nb_flat(plan);
        LAUNCH("k_img_B", k_img_B<<<dim3(nbF, 1, R), block, 0, st>>>(plan->iwe, plan->HW, plan->edges, gNdiv, plan->HW, nbF, plan->sc->ref, 1,
                                                   plan->sc->coefB, plan->part, plan->sc->ref, &plan->sc->counters[1]));
        if (use_tv) {
            const dim3 gridV((W + kTvTX - 1) / kTvTX, (H + kTvTY - 1) / kTvTY), blockV(kTvTX, kTvTY);
            if ((rc = ensure_theta_full(plan, st))) return rc;
            LAUNCH("k_tv", k_tv<<<gridV, blockV, 0, st>>>(plan->theta_full, plan->mask, H, W, gridV.x * gridV.y, plan->Gtv, plan->part, plan->sc));
        }
        LAUNCH("k_scalars(1)", k_scalars<<<1, 32, 0, st>>>(plan->sc, R, (double)plan->HW, hp->alpha, hp->beta, hp->gamma, hp->delta, use_tv, use_div, 1, loss_out));
        if (!want_grad) return EINCM_OK;

        LAUNCH("k_img_C", k_img_C<<<gridT, block, 0, st>>>(plan->iwe, plan->edges, gNdiv, H, W, plan->sc->ref, plan->sc->coefA, plan->sc->coefB, plan->dldi,
                                                               plan->exact ? nullptr : plan->dldi32));
    }
    if (!g_zeroed) CU(cudaMemsetAsync(plan->G, 0, (size_t)plan->HW * 2 * sizeof(double), st));
    if (plan->n_events > 0) {
        const int64_t n = plan->n_stream;
        const int grid = event_grid(plan, n, 256);
        if (plan->exact) {
            if (plan->wrap)
                LAUNCH("k_backward_events", k_backward_events<true><<<grid, 256, 0, st>>>(plan->ev_xy, plan->ev_t, n, plan->theta_full, H, W, R,
                                                                                          plan->tref, plan->dldi, plan->G));
            else
                LAUNCH("k_backward_events", k_backward_events<false><<<grid, 256, 0, st>>>(plan->ev_xy, plan->ev_t, n, plan->theta_full, H, W, R,
                                                                                           plan->tref, plan->dldi, plan->G));
        } else {
            const int gridT2 = std::max(1, plan->n_chunks);
#define BWDT(WR, RB) LAUNCH("k_backward_events", launch_pdl(k_backward_tile<WR, RB>, dim3(gridT2), dim3(256), RB * kWinCap * sizeof(float), st, (const uint32_t*)plan->ev_xy, \
                                   (const double*)plan->ev_t, (const Chunk*)plan->chunks, (const unsigned int*)(plan->totals + 1), plan->tsrc, H, W, R, plan->tref, \
                                   (const float*)plan->dldi32, (const int4*)plan->chunk_win, plan->G, plan->eval_skip))
#define BWDT_RB(WR) do { switch (refs_per_pass(R)) { case 1: BWDT(WR, 1); break; case 2: BWDT(WR, 2); break; \
                                                           case 3: BWDT(WR, 3); break; default: BWDT(WR, 4); } } while (0)
            if (plan->wrap) BWDT_RB(true); else BWDT_RB(false);
#undef BWDT_RB
#undef BWDT
        }
    }
    AxisTaps ty, tx;
    if ((rc = build_axis_taps(plan, h, H, &ty))) return rc;
    if ((rc = build_axis_taps(plan, w, W, &tx))) return rc;
    const double2* Gtv = (use_tv && plan->split_rank == 0) ? plan->Gtv : nullptr;
    const int n_el = h * w;
    const bool handover = plan->last_prev != nullptr;
    if (n_el <= kGatherMaxTiles) {
        // accumulate straight into the caller's gradient (an internal buffer when only d/d alpha is wanted)
        double* gout = grad_out ? grad_out : plan->grad_buf;
        if (!g_zeroed_grad) CU(cudaMemsetAsync(gout, 0, (size_t)n_el * 2 * sizeof(double), st));
        // support of one theta element: <= 2 cells of the resize (+ rounding) per axis
        const int max_ny = std::min(H, 2 * ((H + h - 1) / h) + 2), max_nx = std::min(W, 2 * ((W + w - 1) / w) + 2);
        const int SX = (max_nx + kTgCols - 1) / kTgCols;
        const int SY = (max_ny + kTgTrips * kTgRows - 1) / (kTgTrips * kTgRows);
        const int n_items = n_el * SY * SX;
        const int gridG = (n_items + kTgWarps - 1) / kTgWarps;
        if (Gtv != nullptr)
            LAUNCH("k_theta_grad", k_theta_grad<true><<<gridG, kTgWarps * 32, 0, st>>>(
                (const double2*)plan->G, Gtv, plan->sc, hp->gamma, h, w, H, W, SY, SX, n_items, ty, tx,
                handover ? plan->last_prev : nullptr, plan->last_theta, gout, loss_out, host_out, grad_out != nullptr ? 1 : 0, plan->eval_skip));
        else
            LAUNCH("k_theta_grad", launch_pdl(k_theta_grad<false>, dim3(gridG), dim3(kTgWarps * 32), 0, st,
                (const double2*)plan->G, (const double2*)nullptr, plan->sc, hp->gamma, h, w, H, W, SY, SX, n_items, ty, tx,
                (const double*)(handover ? plan->last_prev : nullptr), plan->last_theta, gout, (const double*)loss_out, host_out, grad_out != nullptr ? 1 : 0,
                plan->eval_skip));
        plan->host_delivered = host_out != nullptr;
    } else {
        CU(cudaMemsetAsync(&plan->sc->dalpha, 0, sizeof(double), st));
        CU(cudaMemsetAsync(plan->grad_buf, 0, (size_t)n_el * 2 * sizeof(double), st));
        const dim3 b2(32, 8), g2((W + 31) / 32, (H + 7) / 8);
        LAUNCH("k_theta_grad_scatter", k_theta_grad_scatter<<<g2, b2, 0, st>>>((const double2*)plan->G, Gtv, plan->sc, hp->gamma, h, w, H, W, ty, tx, plan->grad_buf));
        LAUNCH("k_grad_out", k_grad_out<<<std::min(plan->sm_count * 4, (n_el + 255) / 256), 256, 0, st>>>(nullptr, 0, plan->grad_buf, n_el,
                                                                                   handover ? plan->last_prev : nullptr, plan->last_theta, grad_out, plan->sc));
    }
    if (dalpha_out) CU(cudaMemcpyAsync(dalpha_out, &plan->sc->dalpha, sizeof(double), cudaMemcpyDeviceToDevice, st));
    return EINCM_OK;
}

}  // namespace

extern "C" {

int eincm_abi_version(void) { return EINCM_ABI_VERSION; }

const char* eincm_last_error(const eincm_plan* plan) { return plan ? plan->error.c_str() : g_create_error.c_str(); }

int eincm_plan_create(eincm_plan** out, int device, int H, int W, int64_t max_events, int max_refs, unsigned flags) {
    eincm_plan* plan = nullptr;
    if (!out) return fail(nullptr, EINCM_EINVAL, "out is NULL");
    *out = nullptr;
    if (H < 3 || W < 3 || H > 32767 || W > 32767) return fail(nullptr, EINCM_EINVAL, "sensor size (%d,%d) out of range", H, W);
    if (max_events < 0 || max_events > 0x7fffffffLL) return fail(nullptr, EINCM_EINVAL, "max_events out of range");
    if (max_refs < 1 || max_refs > EINCM_MAX_REFS) return fail(nullptr, EINCM_EINVAL, "max_refs must be in 1..%d", EINCM_MAX_REFS);
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, EINCM_ECUDA, "no CUDA device (%s): this library has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= n_dev) return fail(nullptr, EINCM_EINVAL, "device %d out of range (%d devices)", device, n_dev);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(nullptr, EINCM_ECUDA, "%s", cudaGetErrorString(e));
    if (prop.major != 10) return fail(nullptr, EINCM_ECUDA, "device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, EINCM_ECUDA, "%s", cudaGetErrorString(e));

    {   // the tile kernels keep up to kMaxRB windows of kWinCap cells in dynamic shared memory (> 48 KB needs the opt-in)
        cudaError_t ea = cudaSuccess;
        // every kernel of an evaluation asks for the same shared-memory carve-out (the maximum): the L1 / shared-memory split of an
        // SM can only change while the SM is idle, so kernels of concurrent windows (one stream each) that prefer different splits
        // cannot share an SM and serialise
        auto max_carveout = [&](const void* fn) {
            if (ea == cudaSuccess) ea = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        };
        auto opt_in = [&](const void* fn, int rb) {
            if (ea == cudaSuccess) ea = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, rb * kWinCap * (int)sizeof(uint32_t));
            max_carveout(fn);
        };
#define OPT_IN_RB(RB) opt_in((const void*)k_splat_tile<true, RB>, RB); opt_in((const void*)k_splat_tile<false, RB>, RB); \
                      opt_in((const void*)k_backward_tile<true, RB>, RB); opt_in((const void*)k_backward_tile<false, RB>, RB)
        OPT_IN_RB(1); OPT_IN_RB(2); OPT_IN_RB(3); OPT_IN_RB(4);
#undef OPT_IN_RB
        max_carveout((const void*)k_image_grad);
        max_carveout((const void*)k_theta_grad<false>); max_carveout((const void*)k_theta_grad<true>);
        max_carveout((const void*)k_theta_grad_scatter); max_carveout((const void*)k_tv); max_carveout((const void*)k_upsample_theta);
        max_carveout((const void*)k_image_stats);
        if (ea != cudaSuccess) return fail(nullptr, EINCM_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(ea));
    }
    plan = new (std::nothrow) eincm_plan();
    if (!plan) return fail(nullptr, EINCM_ENOMEM, "host allocation failed");
    plan->device = device; plan->H = H; plan->W = W; plan->HW = (int64_t)H * W; plan->max_events = max_events;
    plan->max_refs = max_refs; plan->flags = flags; plan->wrap = !(flags & EINCM_FLAG_NO_WRAP_NEGATIVE);
    plan->exact = (flags & EINCM_FLAG_EXACT_F64) != 0;
    { const char* gu = std::getenv("EINCM_GRAPH_UNROLL"); if (gu != nullptr) plan->graph_unroll = std::max(0, std::min(64, std::atoi(gu))); }
    plan->sm_count = prop.multiProcessorCount;
    plan->coop_ok = true;            // the streaming image pass has no limit on the sensor width
    plan->tiles_x = (W + kSortTile - 1) / kSortTile;
    plan->n_tiles = plan->tiles_x * ((H + kSortTile - 1) / kSortTile);
    plan->n_keys = plan->n_tiles * kKeysPerTile;
    plan->stream_cap = (max_events + kStreamAlign - 1) / kStreamAlign * kStreamAlign + (int64_t)kStreamAlign * plan->n_tiles + kStreamAlign;
    plan->chunk_cap = (int)(max_events / kChunkEvents) + plan->n_tiles + 1;
    const size_t HW = (size_t)plan->HW, NE = (size_t)max_events, RR = (size_t)max_refs;
    const int nbT = img_tiles_x(plan) * img_tiles_y(plan);
    plan->part_doubles = std::max(10 * max_refs * std::max(nbT, plan->sm_count * 4), image_stats_items(H, W, max_refs) * kFPart + 2 * nbT) + 4096;
    auto body = [&]() -> int {
        // sorted event stream: every tile segment padded to whole groups of kEvK events (k_prep.cuh)
        const size_t NP = (size_t)plan->stream_cap;
        (void)NE;
        CU(dmalloc(&plan->ev_xy, NP)); CU(dmalloc(&plan->ev_t, NP)); CU(dmalloc(&plan->perm, NP));
        CU(dmalloc(&plan->ev_t2, NP)); CU(dmalloc(&plan->perm2, NP));
        CU(cudaMemset(plan->ev_t, 0, NP * sizeof(double))); CU(cudaMemset(plan->ev_t2, 0, NP * sizeof(double)));
        CU(dmalloc(&plan->counts, (size_t)plan->n_keys)); CU(dmalloc(&plan->cursor, (size_t)plan->n_keys));
        CU(dmalloc(&plan->tile_cnt, (size_t)plan->n_tiles + 1)); CU(dmalloc(&plan->tile_start, (size_t)plan->n_tiles + 1));
        CU(dmalloc(&plan->chunk_first, (size_t)plan->n_tiles + 1)); CU(dmalloc(&plan->totals, 4));
        CU(cudaMemset(plan->totals, 0, 4 * sizeof(unsigned int)));
        CU(dmalloc(&plan->chunks, (size_t)plan->chunk_cap));
        CU(dmalloc(&plan->chunks2, (size_t)plan->chunk_cap));
        CU(dmalloc(&plan->chunk_tr, (size_t)plan->chunk_cap));
        CU(dmalloc(&plan->mask, HW));
        CU(dmalloc(&plan->theta_full, HW)); CU(dmalloc(&plan->Gtv, HW));
        CU(dmalloc(&plan->partial, (size_t)kGatherMaxTiles * 2 + (size_t)4 * plan->sm_count));
        CU(dmalloc(&plan->G, HW * 2)); CU(dmalloc(&plan->iwe, RR * HW)); CU(dmalloc(&plan->zero_iwe, HW));
        CU(dmalloc(&plan->dldi, RR * HW)); CU(dmalloc(&plan->edges, RR * HW));
        if (!plan->exact) {
            CU(dmalloc(&plan->dldi32, RR * HW));
            CU(dmalloc(&plan->adj32, RR * HW));
            CU(dmalloc(&plan->iwe_fix, RR * HW));
            CU(dmalloc(&plan->chunk_win, (size_t)plan->chunk_cap * RR));
        }
        CU(dmalloc(&plan->part, (size_t)plan->part_doubles));
        CU(dmalloc(&plan->sc, 1));
        CU(cudaMemset(plan->sc, 0, sizeof(DevScalars)));
        CU(dmalloc(&plan->theta_stage, HW * 2)); CU(dmalloc(&plan->prev_stage, HW * 2));
        CU(dmalloc(&plan->grad_stage, HW * 2 + 8)); CU(dmalloc(&plan->grad_buf, HW * 2));     // + 8: loss slot right behind a staged gradient
        CU(dmalloc(&plan->out_stage, 8));
        CU(cudaHostAlloc((void**)&plan->h_pinned, (HW * 2 + 1024) * sizeof(double), cudaHostAllocMapped));
        // the sequence slot the host entry points poll must not hold a stale number of an earlier plan (freed pinned memory is reused)
        std::memset(plan->h_pinned, 0, (HW * 2 + 1024) * sizeof(double));
        CU(cudaHostGetDevicePointer((void**)&plan->h_mapped_dev, plan->h_pinned, 0));
        CU(cudaStreamCreateWithFlags(&plan->own_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&plan->window_ev, cudaEventDisableTiming));
        if (flags & EINCM_FLAG_BLOCKING_SYNC) CU(cudaEventCreateWithFlags(&plan->wait_ev, cudaEventBlockingSync | cudaEventDisableTiming));
        CU(cudaMallocHost((void**)&plan->h_flag, sizeof(int) * 4));
        return EINCM_OK;
    };
    const int rc = body();
    if (rc != EINCM_OK) {
        g_create_error = plan->error;
        eincm_plan_destroy(plan);
        return rc;
    }
    *out = plan;
    return EINCM_OK;
}

void eincm_plan_destroy(eincm_plan* plan) {
    if (!plan) return;
    cudaSetDevice(plan->device);
    if (plan->window_ev) cudaEventDestroy(plan->window_ev);
    void* bufs[] = {plan->ev_xy, plan->ev_t, plan->perm, plan->ev_t2, plan->perm2, plan->counts, plan->cursor, plan->tile_cnt, plan->tile_start,
                    plan->chunk_first, plan->totals, plan->adj32, plan->chunks, plan->chunks2, plan->chunk_tr, plan->chunk_win, plan->iwe_fix, plan->mask, plan->theta_full,
                    plan->Gtv, plan->partial, plan->G, plan->iwe, plan->zero_iwe, plan->dldi, plan->edges, plan->sbar, plan->gNdiv,
                    plan->part, plan->sc, plan->dldi32, plan->theta_stage, plan->prev_stage, plan->grad_stage, plan->grad_buf, plan->out_stage,
                    plan->xs_stage, plan->ys_stage, plan->ts_stage, plan->edges_stage};
    for (void* b : bufs) if (b) cudaFree(b);
    for (auto& kv : plan->taps_cache) if (kv.second.blob) cudaFree(kv.second.blob);
    for (auto& sp : plan->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto& e : plan->event_pool) cudaEventDestroy(e);
    for (void* q : plan->peer_opened) if (q) cudaIpcCloseMemHandle(q);
    for (auto& kv : plan->bfgs_graphs) {
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
    }
    if (plan->bfgs_state) cudaFree(plan->bfgs_state);
    if (plan->bfgs_vec) cudaFree(plan->bfgs_vec);
    if (plan->bfgs_H) cudaFree(plan->bfgs_H);
    if (plan->own_stream) cudaStreamDestroy(plan->own_stream);
    if (plan->h_pinned) cudaFreeHost(plan->h_pinned);
    if (plan->h_flag) cudaFreeHost(plan->h_flag);
    if (plan->wait_ev) cudaEventDestroy(plan->wait_ev);
    delete plan;
}

int eincm_plan_info(const eincm_plan* plan, int* H, int* W, int64_t* n_events, int* n_refs, int* max_refs) {
    if (!plan) return EINCM_EINVAL;
    if (H) *H = plan->H;
    if (W) *W = plan->W;
    if (n_events) *n_events = plan->n_events;
    if (n_refs) *n_refs = plan->R;
    if (max_refs) *max_refs = plan->max_refs;
    return EINCM_OK;
}

int eincm_plan_set_window_device_ts(eincm_plan* plan, const int16_t* xs, const int16_t* ys, const double* ts, int64_t n, const double* edges,
                                    const double* edge_ts_dev, int R, void* cuda_stream) {
    if (!plan) return EINCM_EINVAL;
    if (R < 1 || R > plan->max_refs) return fail(plan, EINCM_EINVAL, "n_refs %d outside 1..max_refs=%d", R, plan->max_refs);
    if (!edge_ts_dev) return fail(plan, EINCM_EINVAL, "NULL operand");
    CU(cudaSetDevice(plan->device));
    double t_host[EINCM_MAX_REFS];
    cudaStream_t st = (cudaStream_t)cuda_stream;
    // the reference times become kernel constants: one small copy and one synchronisation per window (set_window synchronises anyway)
    CU(cudaMemcpyAsync(t_host, edge_ts_dev, (size_t)R * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return eincm_plan_set_window(plan, xs, ys, ts, n, edges, t_host, R, cuda_stream);
}

int eincm_plan_set_window(eincm_plan* plan, const int16_t* xs, const int16_t* ys, const double* ts, int64_t n, const double* edges,
                          const double* edge_ts_host, int R, void* cuda_stream) {
    if (!plan) return EINCM_EINVAL;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (n < 0 || n > plan->max_events) return fail(plan, EINCM_EINVAL, "n_events %lld exceeds the plan's max_events %lld", (long long)n, (long long)plan->max_events);
    if (R < 1 || R > plan->max_refs) return fail(plan, EINCM_EINVAL, "n_refs %d outside 1..max_refs=%d", R, plan->max_refs);
    if ((n > 0 && (!xs || !ys || !ts)) || !edges || !edge_ts_host) return fail(plan, EINCM_EINVAL, "NULL operand");
    CU(cudaSetDevice(plan->device));
    plan->window_set = false; plan->window_final = false; plan->forward_done = false; plan->zero_div_valid = false;
    plan->fused_pending = false;
    plan->n_events = n; plan->R = R;
    for (int r = 0; r < EINCM_MAX_REFS; ++r) plan->tref.t[r] = r < R ? edge_ts_host[r] : 0.0;

    // multi-reference weights (losses.py:39-46): stats.norm.pdf(linspace(-1.5, 1.5, R)) normalised
    RefTimes wts{};
    {
        double sum = 0.0;
        for (int r = 0; r < R; ++r) {
            // np.linspace(-1.5, 1.5, R): start + r*step (last sample is exactly the stop value)
            double x = R == 1 ? -1.5 : (r == R - 1 ? 1.5 : -1.5 + (double)r * (3.0 / (double)(R - 1)));
            wts.t[r] = std::exp(-0.5 * x * x) / std::sqrt(2.0 * M_PI);
            sum += wts.t[r];
        }
        for (int r = 0; r < R; ++r) wts.t[r] /= sum;
    }
    LAUNCH("k_set_weights", k_set_weights<<<1, 32, 0, st>>>(plan->sc, wts, R));

    CU(cudaMemsetAsync(plan->counts, 0, (size_t)plan->n_keys * sizeof(unsigned int), st));
    CU(cudaMemsetAsync(&plan->sc->error_flag, 0, sizeof(int), st));
    CU(cudaMemsetAsync(plan->sc->counters, 0, sizeof(plan->sc->counters), st));
    if (n > 0) {
        const int grid = event_grid(plan, n, 256);
        LAUNCH("k_histogram", k_histogram<<<grid, 256, 0, st>>>(xs, ys, n, plan->H, plan->W, plan->tiles_x, plan->counts, &plan->sc->error_flag));
    }
    LAUNCH("k_tile_counts", k_tile_counts<<<plan->n_tiles, kKeysPerTile, 0, st>>>(plan->counts, plan->tile_cnt));
    LAUNCH("k_tile_layout", k_tile_layout<<<1, 1024, 0, st>>>(plan->tile_cnt, plan->n_tiles, plan->tile_start, plan->chunk_first, plan->totals));
    LAUNCH("k_tile_finish", k_tile_finish<<<plan->n_tiles, kKeysPerTile, 0, st>>>(plan->counts, plan->tile_cnt, plan->tile_start, plan->chunk_first,
                                                                                  plan->tiles_x, plan->cursor, plan->chunks));
    // until the totals are read back (end of this call) the host uses upper bounds; kernels skip sentinels / read the chunk count on device
    plan->n_stream = std::min<int64_t>(plan->stream_cap, (n + kStreamAlign - 1) / kStreamAlign * kStreamAlign + (int64_t)kStreamAlign * plan->n_tiles);
    plan->n_chunks = (int)std::min<int64_t>(plan->chunk_cap, n / kChunkEvents + plan->n_tiles + 1);
    if (n > 0 && plan->n_chunks <= kMaxOrderedChunks) {          // largest chunks first (see k_chunk_order)
        LAUNCH("k_chunk_order", k_chunk_order<<<(plan->n_chunks + 255) / 256, 256, 0, st>>>(plan->chunks, plan->totals + 1, plan->chunks2));
        std::swap(plan->chunks, plan->chunks2);
    }
    CU(cudaMemsetAsync(plan->ev_xy, 0xff, (size_t)plan->n_stream * sizeof(uint32_t), st));
    LAUNCH("k_event_mask", k_event_mask<<<(int)((plan->HW + 255) / 256), 256, 0, st>>>(plan->counts, plan->H, plan->W, plan->tiles_x, plan->mask));
    if (n > 0) {
        const int grid = event_grid(plan, n, 256);
        LAUNCH("k_scatter_events", k_scatter_events<<<grid, 256, 0, st>>>(xs, ys, ts, n, plan->H, plan->W, plan->tiles_x, plan->cursor, plan->ev_xy, plan->ev_t, plan->perm));
        // stable order inside each pixel (by original index = by time), then the buffers swap roles
        LAUNCH("k_rank_sort_segments", k_rank_sort_segments<<<grid, 256, 0, st>>>(plan->ev_xy, plan->ev_t, plan->perm, plan->n_stream, plan->W, plan->tiles_x,
                                                                                   plan->counts, plan->cursor, plan->ev_t2, plan->perm2));
        std::swap(plan->ev_t, plan->ev_t2);
        std::swap(plan->perm, plan->perm2);
        LAUNCH("k_chunk_trange", k_chunk_trange<<<std::max(1, std::min(plan->n_chunks, plan->sm_count * 8)), 256, 0, st>>>(plan->ev_xy, plan->ev_t, plan->chunks,
                                                                                                          plan->totals + 1, plan->chunk_tr));
    }
    CU(cudaMemcpyAsync(plan->edges, edges, (size_t)R * plan->HW * sizeof(double), cudaMemcpyDeviceToDevice, st));
    LAUNCH("k_edge_sums", k_edge_sums<<<R, 1024, 0, st>>>(plan->edges, plan->HW, plan->sc));
    // zero-warp IWE (losses.py:54): theta = 0 => x' = x for every reference time
    {
        RefTimes z{};
        // event split with peer access: every rank votes into every rank's image; the float64 copy is taken after the
        // cross-rank barrier (eincm_split_window_images)
        int rc = splat_images(plan, ThetaSrc{}, nullptr, 1, z, plan->zero_iwe, "k_splat(zero)", st, plan->n_peers == 0);
        if (rc) return rc;
    }
    // validate (the reference's loaders guarantee in-sensor events; a violation would corrupt the gather at
    // event_warpers.py:34-35, so it is an error here).  One 4-byte read back per window.
    // The zero-warp statistics are launched BEFORE the read-back and its synchronisation: when this call returns, no kernel of the
    // staging is still in flight on `st` (an evaluation on another stream - the plan's own stream - may follow at once).
    if (!(plan->flags & EINCM_FLAG_EVENT_SPLIT)) {
        const int rcf = window_finalize_impl(plan, st);
        if (rcf) return rcf;
    }
    CU(cudaMemcpyAsync(plan->h_flag, &plan->sc->error_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(plan->h_flag + 1, plan->totals, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (plan->h_flag[0] != 0) { plan->window_final = false; return fail(plan, EINCM_ERANGE, "an event lies outside the %dx%d sensor", plan->H, plan->W); }
    plan->n_stream = (int64_t)(unsigned int)plan->h_flag[1];
    plan->n_chunks = (int)(unsigned int)plan->h_flag[2];
    plan->window_set = true;
    ++plan->window_gen;
    plan->window_ev_valid = false;           // everything is complete (synchronised above): nothing to wait for
    plan->window_waited_valid = false;
    return EINCM_OK;
}

int eincm_window_finalize(eincm_plan* plan, void* cuda_stream) {
    if (!plan) return EINCM_EINVAL;
    if (!plan->window_set) return fail(plan, EINCM_ESTATE, "eincm_window_finalize before set_window");
    CU(cudaSetDevice(plan->device));
    return window_finalize_impl(plan, (cudaStream_t)cuda_stream);
}

int eincm_forward_events(eincm_plan* plan, const double* theta, int h, int w, const eincm_hparams* hp, void* cuda_stream) {
    if (!plan) return EINCM_EINVAL;
    CU(cudaSetDevice(plan->device));
    return forward_events_impl(plan, theta, nullptr, 0.0, h, w, hp, (cudaStream_t)cuda_stream);
}

int eincm_backward(eincm_plan* plan, const eincm_hparams* hp, double* loss_out, double* grad_out, void* cuda_stream) {
    if (!plan) return EINCM_EINVAL;
    CU(cudaSetDevice(plan->device));
    return backward_impl(plan, hp, loss_out, grad_out, nullptr, (cudaStream_t)cuda_stream);
}

int eincm_value_and_grad(eincm_plan* plan, const double* theta, int h, int w, const eincm_hparams* hp, double* loss_out,
                         double* grad_out, void* cuda_stream) {
    if (!plan) return EINCM_EINVAL;
    if (plan->flags & EINCM_FLAG_EVENT_SPLIT) return fail(plan, EINCM_ESTATE, "event-split plans use the split-phase calls");
    CU(cudaSetDevice(plan->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc = forward_events_impl(plan, theta, nullptr, 0.0, h, w, hp, st);
    if (rc) return rc;
    return backward_impl(plan, hp, loss_out, grad_out, nullptr, st);
}

int eincm_handover_value_and_grad(eincm_plan* plan, double alpha_handover, const double* prev_theta, const double* theta, int h, int w,
                                  const eincm_hparams* hp, double* loss_out, double* dalpha_out, void* cuda_stream) {
    if (!plan) return EINCM_EINVAL;
    if (plan->flags & EINCM_FLAG_EVENT_SPLIT) return fail(plan, EINCM_ESTATE, "event-split plans use the split-phase calls");
    if (!prev_theta) return fail(plan, EINCM_EINVAL, "prev_theta is NULL");
    CU(cudaSetDevice(plan->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc = forward_events_impl(plan, theta, prev_theta, alpha_handover, h, w, hp, st);
    if (rc) return rc;
    return backward_impl(plan, hp, loss_out, nullptr, dalpha_out, st);
}

}  // extern "C"

namespace {

inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Enqueues one host-operand evaluation on `st`: theta through the pinned staging area, the four kernels, and - when the
// results do not come back through mapped memory - the device -> host copy.  host_collect waits and hands the results over.
int host_enqueue_impl(eincm_plan* plan, const double* theta_host, int h, int w, const eincm_hparams* hp, bool want_grad, cudaStream_t st) {
    if (!theta_host) return fail(plan, EINCM_EINVAL, "NULL operand");
    if (h < 1 || w < 1 || h > plan->H || w > plan->W) return fail(plan, EINCM_EINVAL, "theta shape (%d,%d) outside the sensor", h, w);
    if (plan->flags & EINCM_FLAG_EVENT_SPLIT) return fail(plan, EINCM_ESTATE, "event-split plans use the split-phase calls");
    CU(cudaSetDevice(plan->device));
    const size_t nb = (size_t)h * w * 2 * sizeof(double);
    std::memcpy(plan->h_pinned, theta_host, nb);
    CU(cudaMemcpyAsync(plan->theta_stage, plan->h_pinned, nb, cudaMemcpyHostToDevice, st));
    // results come back through mapped pinned memory written by the last kernel; the staged copy is the fallback for the
    // paths that end in other kernels (dense theta, loss only)
    const size_t n_g = want_grad ? (size_t)h * w * 2 : 0;
    plan->host_ng = n_g;
    double* loss_dev = plan->grad_stage + n_g;
    int rc = forward_events_impl(plan, plan->theta_stage, nullptr, 0.0, h, w, hp, st);
    if (rc) return rc;
    // results come back as [sequence | loss | d alpha | gradient] in the mapped area behind the theta staging area when they fit
    const bool small = n_g + 3 <= 1000;
    rc = backward_impl(plan, hp, loss_dev, want_grad ? plan->grad_stage : nullptr, nullptr, st,
                       small ? plan->h_mapped_dev + (size_t)plan->HW * 2 + 16 : nullptr);
    if (rc) return rc;
    if (plan->host_delivered) plan->host_seq += 1.0;                   // the delivering kernel counts the same way (DevScalars::eval_seq)
    if (!plan->host_delivered)
        CU(cudaMemcpyAsync(plan->h_pinned, plan->grad_stage, (n_g + 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
    return EINCM_OK;
}

// waits for the work enqueued on `st`: spinning (lowest latency) or, with EINCM_FLAG_BLOCKING_SYNC, asleep on an event
int host_wait(eincm_plan* plan, cudaStream_t st) {
    if (plan->wait_ev != nullptr) {
        CU(cudaEventRecord(plan->wait_ev, st));
        CU(cudaEventSynchronize(plan->wait_ev));
    } else {
        CU(cudaStreamSynchronize(st));
    }
    return EINCM_OK;
}

int host_collect_impl(eincm_plan* plan, double* loss_out_host, double* grad_out_host, cudaStream_t st) {
    const double* h_map = plan->h_pinned + (size_t)plan->HW * 2 + 16;   // [sequence | loss | d alpha | gradient] (delivered by the last kernel)
    bool arrived = false;
    if (plan->host_delivered && plan->wait_ev == nullptr) {
        // poll the sequence number in mapped pinned memory (bounded; an error or a very long evaluation ends in the stream
        // synchronisation below)
        const volatile double* seq = h_map;
        // a short pure spin (lowest latency), then yield between polls: on a host with fewer cores than driving threads the
        // optimizers of the other sequences get the core while this thread waits; with idle cores the yield returns at once
        const double t_end = now_s() + 2.0;
        for (long spin = 0;; ++spin) {
            if (*seq == plan->host_seq) { arrived = true; break; }
            if (spin < 4000) {
#if defined(__x86_64__) || defined(__i386__)
                __builtin_ia32_pause();
#endif
            } else {
                std::this_thread::yield();
                if ((spin & 1023) == 0 && now_s() > t_end) break;
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
    }
    if (!arrived) {
        const int rcw = host_wait(plan, st);
        if (rcw) return rcw;
    }
    if (plan->host_delivered) {
        *loss_out_host = h_map[1];
        if (grad_out_host) std::memcpy(grad_out_host, h_map + 3, plan->host_ng * sizeof(double));
    } else {                                                            // staged copy: [gradient | loss]
        *loss_out_host = plan->h_pinned[plan->host_ng];
        if (grad_out_host) std::memcpy(grad_out_host, plan->h_pinned, plan->host_ng * sizeof(double));
    }
    return EINCM_OK;
}

int host_enqueue(eincm_plan* plan, const double* theta_host, int h, int w, const eincm_hparams* hp, bool want_grad, cudaStream_t st) {
    const double t0 = now_s();
    const int rc = host_enqueue_impl(plan, theta_host, h, w, hp, want_grad, st);
    plan->host_enqueue_s += now_s() - t0;
    return rc;
}

int host_collect(eincm_plan* plan, double* loss_out_host, double* grad_out_host, cudaStream_t st) {
    const double t0 = now_s();
    const int rc = host_collect_impl(plan, loss_out_host, grad_out_host, st);
    plan->host_wait_s += now_s() - t0;
    plan->host_evals += 1;
    return rc;
}

}  // namespace

extern "C" {

int eincm_value_and_grad_host(eincm_plan* plan, const double* theta_host, int h, int w, const eincm_hparams* hp, double* loss_out_host,
                              double* grad_out_host, void* cuda_stream) {
    if (!plan) return EINCM_EINVAL;
    if (!loss_out_host) return fail(plan, EINCM_EINVAL, "NULL operand");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int rc = host_enqueue(plan, theta_host, h, w, hp, grad_out_host != nullptr, st);
    if (rc) return rc;
    return host_collect(plan, loss_out_host, grad_out_host, st);
}

}  // extern "C"

// Rendezvous of the objective evaluations of concurrently running minimisations (see include/eincm.h).
struct eincm_group {
    struct Request {
        eincm_plan* plan; const double* x; int h, w; const eincm_hparams* hp; double* f; double* g;
        int rc = 0; bool done = false;
    };
    std::mutex mu;
    std::condition_variable cv;
    int active = 0;                       // members currently inside a minimisation
    bool in_flight = false;               // a burst is being launched / collected
    std::vector<Request*> posted;         // evaluations waiting for the next burst
    int device = -1;
    int burst_percent = 100;              // share of the active members that must have posted before a burst is launched

    void enter() { std::lock_guard<std::mutex> lk(mu); ++active; }
    void leave() {
        { std::lock_guard<std::mutex> lk(mu); --active; }
        cv.notify_all();                  // the remaining members may now be complete: one of the waiters becomes the leader
    }
    // posts one evaluation and returns when it is done (launched by this thread or by another member)
    int evaluate(Request& rq) {
        std::unique_lock<std::mutex> lk(mu);
        posted.push_back(&rq);
        for (;;) {
            if (rq.done) return rq.rc;
            if (!in_flight && !posted.empty() && (int)posted.size() * 100 >= active * burst_percent) {
                // every active member is waiting: this thread launches the burst
                std::vector<Request*> burst;
                burst.swap(posted);
                in_flight = true;
                lk.unlock();
                for (Request* q : burst)
                    q->rc = host_enqueue(q->plan, q->x, q->h, q->w, q->hp, true, q->plan->own_stream);
                for (Request* q : burst) {
                    if (q->rc == EINCM_OK) q->rc = host_collect(q->plan, q->f, q->g, q->plan->own_stream);
                    else cudaStreamSynchronize(q->plan->own_stream);
                }
                lk.lock();
                for (Request* q : burst) q->done = true;
                in_flight = false;
                cv.notify_all();
                continue;                 // rq is part of the burst: returns at the top of the loop
            }
            cv.wait(lk);
        }
    }
};

extern "C" {

int eincm_group_create(eincm_group** out) {
    if (!out) return EINCM_EINVAL;
    *out = new (std::nothrow) eincm_group();
    return *out ? EINCM_OK : EINCM_ENOMEM;
}

void eincm_group_destroy(eincm_group* group) { delete group; }

int eincm_group_set_burst_percent(eincm_group* group, int percent) {
    if (!group || percent < 1 || percent > 100) return EINCM_EINVAL;
    std::lock_guard<std::mutex> lk(group->mu);
    group->burst_percent = percent;
    return EINCM_OK;
}

int eincm_plan_set_group(eincm_plan* plan, eincm_group* group) {
    if (!plan) return EINCM_EINVAL;
    if (group) {
        std::lock_guard<std::mutex> lk(group->mu);
        if (group->device >= 0 && group->device != plan->device) return fail(plan, EINCM_EINVAL, "all plans of a group must live on one device");
        group->device = plan->device;
    }
    plan->group = group;
    return EINCM_OK;
}

int eincm_minimize_bfgs_host(eincm_plan* plan, double* theta_inout_host, int h, int w, const eincm_hparams* hp, int maxiter, double gtol,
                             eincm_opt_result* result_out, void* cuda_stream) {
    if (!plan) return EINCM_EINVAL;
    if (!theta_inout_host || !result_out) return fail(plan, EINCM_EINVAL, "NULL operand");
    if (maxiter < 0) return fail(plan, EINCM_EINVAL, "maxiter must be >= 0");
    if (h < 1 || w < 1 || h > plan->H || w > plan->W) return fail(plan, EINCM_EINVAL, "theta shape (%d,%d) outside the sensor", h, w);
    // BFGS keeps a dense n x n inverse Hessian on the host: tile theta only (the reference solves 1x1 .. 16x16 with it, main.yaml:25-59)
    if ((int64_t)h * w > kGatherMaxTiles)
        return fail(plan, EINCM_EINVAL, "eincm_minimize_bfgs_host: theta %dx%d needs a dense %lld^2 inverse Hessian; the limit is %d elements",
                    h, w, (long long)h * w * 2, kGatherMaxTiles);
    cudaStream_t st = cuda_stream == (void*)(intptr_t)-1 ? plan->own_stream : (cudaStream_t)cuda_stream;
    const int n = h * w * 2;
    eincm_group* grp = plan->group;
    eincm_opt::Objective fun = [&](const double* x, double* f, double* g) -> int {
        if (grp != nullptr) {             // rendezvous with the other sequences of the group: launched in a burst by one thread
            eincm_group::Request rq{plan, x, h, w, hp, f, g};
            return grp->evaluate(rq);
        }
        const int rc = host_enqueue(plan, x, h, w, hp, true, st);
        if (rc) return rc;
        return host_collect(plan, f, g, st);
    };
    int err = 0;
    eincm_opt::Result r;
    if (grp != nullptr) grp->enter();
    // no C++ exception may cross the C boundary (std::bad_alloc of the optimizer's work space, std::function)
    try {
        r = eincm_opt::bfgs(fun, n, theta_inout_host, maxiter, gtol, &err);
    } catch (const std::bad_alloc&) {
        if (grp != nullptr) grp->leave();
        return fail(plan, EINCM_ENOMEM, "eincm_minimize_bfgs_host: out of host memory for n = %d", n);
    } catch (const std::exception& e) {
        if (grp != nullptr) grp->leave();
        return fail(plan, EINCM_ECUDA, "eincm_minimize_bfgs_host: %s", e.what());
    }
    if (grp != nullptr) grp->leave();
    if (err) return err;
    result_out->fun = r.fun; result_out->nit = r.nit; result_out->nfev = r.nfev; result_out->status = r.status; result_out->reserved = 0;
    return EINCM_OK;
}

int eincm_minimize_handover_host(eincm_plan* plan, double* alpha_inout_host, double lo, double hi, const double* prev_theta_host,
                                 const double* theta_host, int h, int w, const eincm_hparams* hp, int maxiter, double pgtol,
                                 eincm_opt_result* result_out, void* cuda_stream) {
    if (!plan) return EINCM_EINVAL;
    if (!alpha_inout_host || !prev_theta_host || !theta_host || !result_out) return fail(plan, EINCM_EINVAL, "NULL operand");
    if (h < 1 || w < 1 || h > plan->H || w > plan->W) return fail(plan, EINCM_EINVAL, "theta shape (%d,%d) outside the sensor", h, w);
    if (!(lo <= hi)) return fail(plan, EINCM_EINVAL, "empty interval [%g, %g]", lo, hi);
    if (plan->flags & EINCM_FLAG_EVENT_SPLIT) return fail(plan, EINCM_ESTATE, "event-split plans use the split-phase calls");
    CU(cudaSetDevice(plan->device));
    cudaStream_t st = cuda_stream == (void*)(intptr_t)-1 ? plan->own_stream : (cudaStream_t)cuda_stream;
    const size_t nb = (size_t)h * w * 2 * sizeof(double);
    // the two flow fields are constant during the solve: stage them once (pageable sources: the copies are synchronous)
    CU(cudaMemcpyAsync(plan->theta_stage, theta_host, nb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(plan->prev_stage, prev_theta_host, nb, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    double* h_out = plan->h_pinned + (size_t)plan->HW * 2;
    eincm_opt::Objective fun = [&](const double* a, double* f, double* g) -> int {
        const int rc = eincm_handover_value_and_grad(plan, *a, plan->prev_stage, plan->theta_stage, h, w, hp, plan->out_stage, plan->out_stage + 1, st);
        if (rc) return rc;
        if (cudaMemcpyAsync(h_out, plan->out_stage, 2 * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess || host_wait(plan, st) != EINCM_OK)
            return fail(plan, EINCM_ECUDA, "copy of the handover result failed");
        *f = h_out[0]; *g = h_out[1];
        return EINCM_OK;
    };
    int err = 0;
    eincm_opt::Result r;
    try {
        r = eincm_opt::bounded_scalar(fun, alpha_inout_host, lo, hi, maxiter, pgtol, 1e7, &err);
    } catch (const std::exception& e) {
        return fail(plan, EINCM_ECUDA, "eincm_minimize_handover_host: %s", e.what());
    }
    if (err) return err;
    result_out->fun = r.fun; result_out->nit = r.nit; result_out->nfev = r.nfev; result_out->status = r.status; result_out->reserved = 0;
    return EINCM_OK;
}

namespace {

// Builds the graph of one level solve: k_bfgs_init -> WHILE { evaluation at x_trial ; k_bfgs_step } -> result to pinned host memory.
int build_level_graph(eincm_plan* plan, eincm_plan::LevelGraph& lg, int h, int w, const eincm_hparams* hp, int maxiter, double gtol,
                      cudaStream_t st, bool plain_launches, int unroll) {
    const int n = h * w * 2;
    double* v = plan->bfgs_vec;
    BfgsBufs B{};
    B.x = v; B.g = v + n; B.p = v + 2 * n; B.s = v + 3 * n; B.y = v + 4 * n; B.Hy = v + 5 * n; B.x_trial = v + 6 * n;
    double* g_trial = v + 7 * n;
    double* f_trial = v + 8 * n;
    B.g_trial = g_trial; B.f_trial = f_trial; B.result = v + 8 * n + 8; B.H = plan->bfgs_H;
    cudaGraph_t graph = nullptr;
    if (unroll > 0) {
        // Unrolled form: `unroll` x { evaluation ; k_bfgs_step } + the result copy, captured as an ordinary graph; k_bfgs_init runs before
        // the first launch and the host relaunches until the record says done.  Once the level has ended, the remaining evaluations and
        // steps of a launch return at once (skip flag).  Ordinary kernel nodes: graphs of different plans overlap like streams do.
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
            return fail(plan, EINCM_ECUDA, "cudaStreamBeginCapture: %s", cudaGetErrorString(cudaGetLastError()));
        t_plain_launches = plain_launches;
        plan->eval_skip = &plan->bfgs_state->done;
        int rcu = EINCM_OK;
        for (int k = 0; k < unroll && rcu == EINCM_OK; ++k) {
            rcu = forward_events_impl(plan, B.x_trial, nullptr, 0.0, h, w, hp, st);
            if (rcu == EINCM_OK) rcu = backward_impl(plan, hp, f_trial, g_trial, nullptr, st);
            if (rcu == EINCM_OK) {
                k_bfgs_step<<<1, kOptNT, 0, st>>>(plan->bfgs_state, B, (cudaGraphConditionalHandle)0, 0);
                if (cudaGetLastError() != cudaSuccess) rcu = fail(plan, EINCM_ECUDA, "launch k_bfgs_step");
            }
        }
        plan->eval_skip = nullptr;
        t_plain_launches = false;
        if (rcu == EINCM_OK && cudaMemcpyAsync(plan->h_pinned, B.result, (size_t)(5 + n) * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess)
            rcu = fail(plan, EINCM_ECUDA, "capture of the result copy");
        const cudaError_t ecu = cudaStreamEndCapture(st, &graph);
        if (rcu != EINCM_OK) { if (graph) cudaGraphDestroy(graph); return rcu; }
        if (ecu != cudaSuccess) { cudaGetLastError(); if (graph) cudaGraphDestroy(graph); return fail(plan, EINCM_ECUDA, "capture of the unrolled level: %s", cudaGetErrorString(ecu)); }
        cudaGraphExec_t exec_u = nullptr;
        const cudaError_t eiu = cudaGraphInstantiate(&exec_u, graph, 0);
        if (eiu != cudaSuccess) { cudaGetLastError(); cudaGraphDestroy(graph); return fail(plan, EINCM_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(eiu)); }
        if (lg.exec) cudaGraphExecDestroy(lg.exec);
        if (lg.graph) cudaGraphDestroy(lg.graph);
        lg.graph = graph; lg.exec = exec_u; lg.gen = plan->window_gen;
        return EINCM_OK;
    }
    CU(cudaGraphCreate(&graph, 0));
    auto bail = [&](int rc) { cudaGraphDestroy(graph); return rc; };
    // node 1: state
    cudaGraphNode_t n_init = nullptr;
    {
        BfgsDev* S = plan->bfgs_state;
        int nn = n, mi = maxiter;
        double gt = gtol;
        void* args[] = {&S, &nn, &mi, &gt};
        cudaKernelNodeParams kp{};
        kp.func = (void*)k_bfgs_init; kp.gridDim = dim3(1); kp.blockDim = dim3(32); kp.sharedMemBytes = 0; kp.kernelParams = args; kp.extra = nullptr;
        if (cudaGraphAddKernelNode(&n_init, graph, nullptr, 0, &kp) != cudaSuccess) return bail(fail(plan, EINCM_ECUDA, "cudaGraphAddKernelNode: %s", cudaGetErrorString(cudaGetLastError())));
    }
    // node 2: the loop
    cudaGraphConditionalHandle handle;
    if (cudaGraphConditionalHandleCreate(&handle, graph, 1u, cudaGraphCondAssignDefault) != cudaSuccess)
        return bail(fail(plan, EINCM_ECUDA, "cudaGraphConditionalHandleCreate: %s", cudaGetErrorString(cudaGetLastError())));
    cudaGraphNodeParams cp{};
    cp.type = cudaGraphNodeTypeConditional;
    cp.conditional.handle = handle; cp.conditional.type = cudaGraphCondTypeWhile; cp.conditional.size = 1;
    cudaGraphNode_t n_loop = nullptr;
    if (cudaGraphAddNode(&n_loop, graph, &n_init, 1, &cp) != cudaSuccess)
        return bail(fail(plan, EINCM_ECUDA, "conditional graph node: %s", cudaGetErrorString(cudaGetLastError())));
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    // body: the evaluation exactly as eincm_value_and_grad enqueues it (device operands), then the optimizer step
    if (cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
        return bail(fail(plan, EINCM_ECUDA, "cudaStreamBeginCaptureToGraph: %s", cudaGetErrorString(cudaGetLastError())));
    t_plain_launches = plain_launches;
    int rc = forward_events_impl(plan, B.x_trial, nullptr, 0.0, h, w, hp, st);
    if (rc == EINCM_OK) rc = backward_impl(plan, hp, f_trial, g_trial, nullptr, st);
    t_plain_launches = false;
    if (rc == EINCM_OK) {
        k_bfgs_step<<<1, kOptNT, 0, st>>>(plan->bfgs_state, B, handle, 1);
        if (cudaGetLastError() != cudaSuccess) rc = fail(plan, EINCM_ECUDA, "launch k_bfgs_step");
    }
    cudaGraph_t captured = nullptr;
    const cudaError_t ec = cudaStreamEndCapture(st, &captured);
    if (rc != EINCM_OK) return bail(rc);
    if (ec != cudaSuccess) { cudaGetLastError(); return bail(fail(plan, EINCM_ECUDA, "capture of the loop body: %s", cudaGetErrorString(ec))); }
    // node 3: result record [fun, nit, nfev, status, x] -> pinned host memory
    cudaGraphNode_t n_out = nullptr;
    if (cudaGraphAddMemcpyNode1D(&n_out, graph, &n_loop, 1, plan->h_pinned, B.result, (size_t)(5 + n) * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
        return bail(fail(plan, EINCM_ECUDA, "cudaGraphAddMemcpyNode1D: %s", cudaGetErrorString(cudaGetLastError())));
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
    if (ei != cudaSuccess) { cudaGetLastError(); return bail(fail(plan, EINCM_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ei))); }
    if (lg.exec) cudaGraphExecDestroy(lg.exec);
    if (lg.graph) cudaGraphDestroy(lg.graph);
    lg.graph = graph; lg.exec = exec; lg.gen = plan->window_gen;
    return EINCM_OK;
}

}  // namespace

namespace {
int minimize_bfgs_graph_impl(eincm_plan* plan, double* theta_inout_host, int h, int w, const eincm_hparams* hp, int maxiter, double gtol,
                             eincm_opt_result* result_out, void* cuda_stream);
}

int eincm_minimize_bfgs_graph_host(eincm_plan* plan, double* theta_inout_host, int h, int w, const eincm_hparams* hp, int maxiter, double gtol,
                                   eincm_opt_result* result_out, void* cuda_stream) {
    if (!plan) return EINCM_EINVAL;
    try {                                    // no C++ exception crosses the C boundary (the graph table allocates)
        return minimize_bfgs_graph_impl(plan, theta_inout_host, h, w, hp, maxiter, gtol, result_out, cuda_stream);
    } catch (const std::bad_alloc&) {
        return fail(plan, EINCM_ENOMEM, "eincm_minimize_bfgs_graph_host: out of host memory");
    } catch (const std::exception& e) {
        return fail(plan, EINCM_ECUDA, "eincm_minimize_bfgs_graph_host: %s", e.what());
    }
}

namespace {
int minimize_bfgs_graph_impl(eincm_plan* plan, double* theta_inout_host, int h, int w, const eincm_hparams* hp, int maxiter, double gtol,
                             eincm_opt_result* result_out, void* cuda_stream) {
    if (!theta_inout_host || !result_out || !hp) return fail(plan, EINCM_EINVAL, "NULL operand");
    if (maxiter < 0) return fail(plan, EINCM_EINVAL, "maxiter must be >= 0");
    if (h < 1 || w < 1 || h > plan->H || w > plan->W) return fail(plan, EINCM_EINVAL, "theta shape (%d,%d) outside the sensor", h, w);
    const int n = h * w * 2;
    if (n > kOptMaxN) return fail(plan, EINCM_EINVAL, "the device-side loop holds up to %d flow parameters (theta %dx%d has %d): use eincm_minimize_bfgs_host", kOptMaxN, h, w, n);
    if (plan->flags & EINCM_FLAG_EVENT_SPLIT) return fail(plan, EINCM_ESTATE, "event-split plans use the split-phase calls");
    if (plan->exact || !plan->coop_ok || hp->delta != 0.0) return fail(plan, EINCM_ESTATE, "the device-side loop runs the default (fused) evaluation path only");
    if (!plan->window_set) return fail(plan, EINCM_ESTATE, "minimize before set_window");
    int rc = check_hp(plan, hp);
    if (rc) return rc;
    CU(cudaSetDevice(plan->device));
    cudaStream_t st = (cuda_stream == (void*)(intptr_t)-1 || cuda_stream == nullptr) ? plan->own_stream : (cudaStream_t)cuda_stream;   // never the legacy stream: it cannot be captured
    if (plan->bfgs_cap_n < n) {
        if (plan->bfgs_vec) { CU(cudaStreamSynchronize(st)); cudaFree(plan->bfgs_vec); cudaFree(plan->bfgs_H); plan->bfgs_vec = nullptr; plan->bfgs_H = nullptr; }
        for (auto& kv : plan->bfgs_graphs) { if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec); if (kv.second.graph) cudaGraphDestroy(kv.second.graph); }
        plan->bfgs_graphs.clear();
        const int cap = std::max(n, 512);
        CU(dmalloc(&plan->bfgs_vec, (size_t)9 * cap + 16 + 8 + cap));
        CU(dmalloc(&plan->bfgs_H, (size_t)cap * cap));
        if (!plan->bfgs_state) CU(dmalloc(&plan->bfgs_state, 1));
        plan->bfgs_cap_n = cap;
    }
    // the unrolled form needs the skip flag in every kernel of the evaluation: the TV kernels (gamma != 0 at level 0) do not take one
    const bool use_tv = hp->gamma != 0.0 && hp->cur_pyr_lvl <= 0;
    const int unroll = use_tv ? 0 : plan->graph_unroll;
    // the WHILE form bakes maxiter and gtol into its init node; the unrolled form takes them at run time
    const std::vector<double> key = {(double)h, (double)w, unroll > 0 ? -1.0 : (double)maxiter, unroll > 0 ? -1.0 : gtol, hp->alpha, hp->beta, hp->gamma,
                                     hp->delta, (double)hp->cur_pyr_lvl, (double)hp->n_pyr_lvls, (double)hp->method, (double)unroll};
    eincm_plan::LevelGraph& lg = plan->bfgs_graphs[key];
    if (lg.exec == nullptr || lg.gen != plan->window_gen) {
        // everything the captured calls would do synchronously or across streams happens here, outside the capture
        AxisTaps ty, tx;
        if ((rc = build_axis_taps(plan, h, plan->H, &ty))) return rc;
        if ((rc = build_axis_taps(plan, w, plan->W, &tx))) return rc;
        if (plan->window_ev_valid && st != plan->window_stream && !(plan->window_waited_valid && plan->window_waited == st)) {
            CU(cudaStreamWaitEvent(st, plan->window_ev, 0));
            plan->window_waited = st; plan->window_waited_valid = true;
        }
        if (!plan->fix_clean) {
            CU(cudaMemsetAsync(plan->iwe_fix, 0, (size_t)plan->max_refs * plan->HW * sizeof(unsigned long long), st));
            plan->fix_clean = true;
        }
        const bool was_timing = plan->timing;
        plan->timing = false;                        // event records are not part of a loop body
        rc = build_level_graph(plan, lg, h, w, hp, maxiter, gtol, st, plan->graph_no_pdl, unroll);
        if (rc != EINCM_OK && !plan->graph_no_pdl) {                     // once more without programmatic edges in the body
            plan->graph_no_pdl = true;
            plan->fix_clean = true;
            rc = build_level_graph(plan, lg, h, w, hp, maxiter, gtol, st, true, unroll);
        }
        plan->timing = was_timing;
        if (rc != EINCM_OK) return rc;
    }
    // run: theta in through the pinned staging area, the whole level on the device, one synchronisation
    const size_t nb = (size_t)n * sizeof(double);
    double* stage = plan->h_pinned + 8 + n;                          // behind the result record [done, fun, nit, nfev, status, theta]
    std::memcpy(stage, theta_inout_host, nb);
    CU(cudaMemcpyAsync(plan->bfgs_vec + 6 * (size_t)n, stage, nb, cudaMemcpyHostToDevice, st));
    if (unroll > 0) {
        k_bfgs_init<<<1, 32, 0, st>>>(plan->bfgs_state, n, maxiter, gtol);
        if (cudaGetLastError() != cudaSuccess) return fail(plan, EINCM_ECUDA, "launch k_bfgs_init");
        // every launch runs up to `unroll` evaluations; a line search takes at most 100 + 21 of them per iteration
        const long max_launches = ((long)maxiter + 1) * 122 / unroll + 2;
        plan->h_pinned[0] = 0.0;
        for (long l = 0;; ++l) {
            CU(cudaGraphLaunch(lg.exec, st));
            rc = host_wait(plan, st);
            if (rc) return rc;
            if (plan->h_pinned[0] != 0.0) break;
            if (l >= max_launches) return fail(plan, EINCM_ECUDA, "the device-side loop did not end within %ld launches", max_launches);
        }
    } else {
        CU(cudaGraphLaunch(lg.exec, st));
        rc = host_wait(plan, st);
        if (rc) return rc;
    }
    // the graph ran the evaluation at least once: the plan's bookkeeping is the one the captured calls left behind
    plan->fix_clean = true; plan->forward_done = true; plan->host_delivered = false;
    const double* r = plan->h_pinned;
    result_out->fun = r[1]; result_out->nit = (int32_t)r[2]; result_out->nfev = (int32_t)r[3]; result_out->status = (int32_t)r[4]; result_out->reserved = 0;
    std::memcpy(theta_inout_host, r + 5, nb);
    plan->host_evals += result_out->nfev;
    return EINCM_OK;
}
}  // namespace

int eincm_value_and_grad_host_batch(eincm_plan* const* plans, int n_plans, const double* const* thetas_host, int h, int w,
                                    const eincm_hparams* hp, double* losses_out_host, double* const* grads_out_host) {
    if (!plans || n_plans < 1 || !plans[0]) return EINCM_EINVAL;
    eincm_plan* plan = plans[0];                                       // errors of the batch are reported on the first plan
    if (!thetas_host || !losses_out_host) return fail(plan, EINCM_EINVAL, "NULL operand");
    for (int k = 0; k < n_plans; ++k) {
        if (!plans[k] || !thetas_host[k]) return fail(plan, EINCM_EINVAL, "plan / theta %d is NULL", k);
        if (plans[k]->device != plan->device) return fail(plan, EINCM_EINVAL, "all plans of a batch must live on one device");
        for (int q = 0; q < k; ++q) if (plans[q] == plans[k]) return fail(plan, EINCM_EINVAL, "plan %d appears twice in the batch", k);
    }
    for (int k = 0; k < n_plans; ++k) {
        const bool want_grad = grads_out_host != nullptr && grads_out_host[k] != nullptr;
        const int rc = host_enqueue(plans[k], thetas_host[k], h, w, hp, want_grad, plans[k]->own_stream);
        if (rc) {
            if (plans[k] != plan) plan->error = plans[k]->error;
            for (int q = 0; q < k; ++q) cudaStreamSynchronize(plans[q]->own_stream);
            return rc;
        }
    }
    int rc_all = EINCM_OK;
    for (int k = 0; k < n_plans; ++k) {
        const int rc = host_collect(plans[k], losses_out_host + k, grads_out_host ? grads_out_host[k] : nullptr, plans[k]->own_stream);
        if (rc && !rc_all) { rc_all = rc; if (plans[k] != plan) plan->error = plans[k]->error; }
    }
    return rc_all;
}

int eincm_handover_value_and_grad_host(eincm_plan* plan, double alpha_handover, const double* prev_theta_host, const double* theta_host,
                                       int h, int w, const eincm_hparams* hp, double* loss_out_host, double* dalpha_out_host,
                                       void* cuda_stream) {
    if (!plan) return EINCM_EINVAL;
    if (!theta_host || !prev_theta_host || !loss_out_host) return fail(plan, EINCM_EINVAL, "NULL operand");
    if (h < 1 || w < 1 || h > plan->H || w > plan->W) return fail(plan, EINCM_EINVAL, "theta shape (%d,%d) outside the sensor", h, w);
    CU(cudaSetDevice(plan->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t nb = (size_t)h * w * 2 * sizeof(double);
    // both operands are staged through the one pinned buffer; the first copy must land before it is reused
    std::memcpy(plan->h_pinned, theta_host, nb);
    CU(cudaMemcpyAsync(plan->theta_stage, plan->h_pinned, nb, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    std::memcpy(plan->h_pinned, prev_theta_host, nb);
    CU(cudaMemcpyAsync(plan->prev_stage, plan->h_pinned, nb, cudaMemcpyHostToDevice, st));
    int rc = eincm_handover_value_and_grad(plan, alpha_handover, plan->prev_stage, plan->theta_stage, h, w, hp, plan->out_stage,
                                           dalpha_out_host ? plan->out_stage + 1 : nullptr, st);
    if (rc) return rc;
    double* h_out = plan->h_pinned + (size_t)plan->HW * 2;
    CU(cudaMemcpyAsync(h_out, plan->out_stage, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *loss_out_host = h_out[0];
    if (dalpha_out_host) *dalpha_out_host = h_out[1];
    return EINCM_OK;
}

int eincm_value_and_grad_stateless_host(eincm_plan* plan, const double* theta_host, int h, int w, const int16_t* xs_host,
                                        const int16_t* ys_host, const double* ts_host, int64_t n, const double* edges_host,
                                        const double* edge_ts_host, int R, const eincm_hparams* hp, double* loss_out_host,
                                        double* grad_out_host, void* cuda_stream) {
    if (!plan) return EINCM_EINVAL;
    if (n < 0 || n > plan->max_events) return fail(plan, EINCM_EINVAL, "n_events exceeds the plan's max_events");
    if (R < 1 || R > plan->max_refs) return fail(plan, EINCM_EINVAL, "n_refs %d outside 1..max_refs=%d", R, plan->max_refs);
    if ((n > 0 && (!xs_host || !ys_host || !ts_host)) || !edges_host || !edge_ts_host) return fail(plan, EINCM_EINVAL, "NULL operand");
    CU(cudaSetDevice(plan->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (!plan->xs_stage) {
        CU(dmalloc(&plan->xs_stage, (size_t)plan->max_events)); CU(dmalloc(&plan->ys_stage, (size_t)plan->max_events));
        CU(dmalloc(&plan->ts_stage, (size_t)plan->max_events));
        CU(dmalloc(&plan->edges_stage, (size_t)plan->max_refs * plan->HW));
    }
    // pageable sources: these copies are synchronous with respect to the host, which is the contract of a *_host call
    CU(cudaMemcpyAsync(plan->xs_stage, xs_host, (size_t)n * sizeof(int16_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(plan->ys_stage, ys_host, (size_t)n * sizeof(int16_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(plan->ts_stage, ts_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(plan->edges_stage, edges_host, (size_t)R * plan->HW * sizeof(double), cudaMemcpyHostToDevice, st));
    int rc = eincm_plan_set_window(plan, plan->xs_stage, plan->ys_stage, plan->ts_stage, n, plan->edges_stage, edge_ts_host, R, st);
    if (rc) return rc;
    return eincm_value_and_grad_host(plan, theta_host, h, w, hp, loss_out_host, grad_out_host, st);
}

double* eincm_zero_iwe_ptr(eincm_plan* plan) { return plan ? plan->zero_iwe : nullptr; }

// debug taps: the fused path keeps the images of the last evaluation as cell records - unpack them (synchronous)
double* eincm_iwe_ptr(eincm_plan* plan) {
    return plan ? plan->iwe : nullptr;
}
uint8_t* eincm_mask_ptr(eincm_plan* plan) { return plan ? plan->mask : nullptr; }
double* eincm_dldi_ptr(eincm_plan* plan) {
    if (!plan) return nullptr;
    if (plan->dldi_stale) {
        // the fused image pass only writes the float32 copy the event kernels read: rebuild the float64 image of the last evaluation
        // from the same operands (debug tap: synchronous)
        if (cudaSetDevice(plan->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return nullptr;
        ImageGradArgs ga{};
        ga.fix = nullptr; ga.edges = plan->edges; ga.iwe = plan->iwe; ga.adj32 = plan->adj32; ga.sc = plan->sc;
        ga.dldi = plan->dldi; ga.dldi32 = nullptr; ga.HW = (int)plan->HW; ga.R = plan->R; ga.want_grad = 1;
        k_image_grad<<<std::max(1, std::min((int)((plan->HW + 1023) / 1024), plan->sm_count * 8)), 256>>>(ga);
        if (cudaDeviceSynchronize() != cudaSuccess) return nullptr;
        plan->dldi_stale = false;
    }
    return plan->dldi;
}
double* eincm_theta_full_ptr(eincm_plan* plan) {
    if (!plan) return nullptr;
    // debug tap: the default path does not materialise the dense field - do it now (synchronous, legacy default stream)
    if (plan->forward_done && !plan->theta_full_valid) {
        cudaSetDevice(plan->device);
        cudaDeviceSynchronize();
        if (ensure_theta_full(plan, (cudaStream_t)0) != EINCM_OK) return nullptr;
        cudaDeviceSynchronize();
    }
    return (double*)plan->theta_full;
}

int eincm_get_scalars(eincm_plan* plan, double* out_host, int n_doubles, void* cuda_stream) {
    if (!plan || !out_host) return EINCM_EINVAL;
    const int need = EINCM_S_HEADER + 5 * plan->max_refs;
    if (n_doubles < need) return fail(plan, EINCM_EINVAL, "out_host needs %d doubles", need);
    CU(cudaSetDevice(plan->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    DevScalars* hs = (DevScalars*)plan->h_pinned;
    static_assert(sizeof(DevScalars) <= 1024 * sizeof(double), "scalar block fits the pinned staging area");
    CU(cudaMemcpyAsync(hs, plan->sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    out_host[EINCM_S_FINAL_LOSS] = hs->loss;
    out_host[EINCM_S_MEAN_REL_CORR] = hs->mean_rel_corr;
    out_host[EINCM_S_MEAN_REL_CONTRAST] = hs->mean_rel_contrast;
    out_host[EINCM_S_MEAN_REL_IWE_DIV] = hs->mean_rel_div;
    out_host[EINCM_S_THETA_TV] = hs->tv;
    out_host[EINCM_S_ZERO_CONTRAST] = hs->zero[0].contrast;
    out_host[EINCM_S_ZERO_IWE_DIV] = plan->zero_div_valid ? hs->zero[0].div : 0.0;
    out_host[EINCM_S_DALPHA_HANDOVER] = hs->dalpha;
    const int M = plan->max_refs;
    for (int r = 0; r < M; ++r) {
        const bool live = r < plan->R;
        out_host[EINCM_S_PER_REF + 0 * M + r] = live ? hs->ref[r].contrast : 0.0;
        out_host[EINCM_S_PER_REF + 1 * M + r] = live ? -hs->ref[r].mse : 0.0;
        out_host[EINCM_S_PER_REF + 2 * M + r] = live ? -hs->zero[r].mse : 0.0;
        out_host[EINCM_S_PER_REF + 3 * M + r] = live ? hs->ref[r].div : 0.0;
        out_host[EINCM_S_PER_REF + 4 * M + r] = live ? hs->weights[r] : 0.0;
    }
    return EINCM_OK;
}

int eincm_plan_set_event_split(eincm_plan* plan, int rank, int world) {
    if (!plan) return EINCM_EINVAL;
    if (!(plan->flags & EINCM_FLAG_EVENT_SPLIT)) return fail(plan, EINCM_ESTATE, "plan was not created with EINCM_FLAG_EVENT_SPLIT");
    if (world < 1 || rank < 0 || rank >= world) return fail(plan, EINCM_EINVAL, "rank %d outside 0..%d", rank, world - 1);
    plan->split_rank = rank; plan->split_world = world;
    return EINCM_OK;
}

int eincm_plan_set_split_fixed_point(eincm_plan* plan, int on) {
    if (!plan) return EINCM_EINVAL;
    if (!(plan->flags & EINCM_FLAG_EVENT_SPLIT)) return fail(plan, EINCM_ESTATE, "plan was not created with EINCM_FLAG_EVENT_SPLIT");
    if (plan->exact || !plan->iwe_fix) return fail(plan, EINCM_ESTATE, "the fixed-point all-reduce needs the default (fixed-point) path");
    plan->split_fixed = on != 0;
    return EINCM_OK;
}

int eincm_plan_ipc_handle(eincm_plan* plan, void* handle_out, int handle_bytes) {
    if (!plan || !handle_out) return EINCM_EINVAL;
    if (plan->exact || !plan->iwe_fix) return fail(plan, EINCM_ESTATE, "peer access needs the default (fixed-point) path");
    if (handle_bytes < (int)sizeof(cudaIpcMemHandle_t)) return fail(plan, EINCM_EINVAL, "handle_out needs %d bytes", (int)sizeof(cudaIpcMemHandle_t));
    CU(cudaSetDevice(plan->device));
    cudaIpcMemHandle_t hnd;
    CU(cudaIpcGetMemHandle(&hnd, plan->iwe_fix));
    std::memcpy(handle_out, &hnd, sizeof hnd);
    return EINCM_OK;
}

namespace {
int check_split_peers(eincm_plan* plan, int n) {
    if (!(plan->flags & EINCM_FLAG_EVENT_SPLIT)) return fail(plan, EINCM_ESTATE, "plan was not created with EINCM_FLAG_EVENT_SPLIT");
    if (plan->exact || !plan->iwe_fix) return fail(plan, EINCM_ESTATE, "peer access needs the default (fixed-point) path");
    if (n != plan->split_world || n < 1 || n > kMaxPeers)
        return fail(plan, EINCM_EINVAL, "expected one entry per rank of the split (%d, at most %d), got %d", plan->split_world, kMaxPeers, n);
    return EINCM_OK;
}
}  // namespace

int eincm_plan_set_peers(eincm_plan* plan, const void* handles, int n_handles) {
    if (!plan || !handles) return EINCM_EINVAL;
    int rc = check_split_peers(plan, n_handles);
    if (rc) return rc;
    CU(cudaSetDevice(plan->device));
    for (int q = 0; q < n_handles; ++q) {
        if (q == plan->split_rank) { plan->peer_fix[q] = plan->iwe_fix; continue; }
        cudaIpcMemHandle_t hnd;
        std::memcpy(&hnd, (const char*)handles + (size_t)q * sizeof hnd, sizeof hnd);
        void* ptr = nullptr;
        CU(cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
        plan->peer_opened[q] = ptr;
        plan->peer_fix[q] = (unsigned long long*)ptr;
    }
    plan->n_peers = n_handles;
    return EINCM_OK;
}

int eincm_plan_set_peer_pointers(eincm_plan* plan, void* const* fix_ptrs, int n_ptrs) {
    if (!plan || !fix_ptrs) return EINCM_EINVAL;
    int rc = check_split_peers(plan, n_ptrs);
    if (rc) return rc;
    for (int q = 0; q < n_ptrs; ++q) plan->peer_fix[q] = (q == plan->split_rank) ? plan->iwe_fix : (unsigned long long*)fix_ptrs[q];
    plan->n_peers = n_ptrs;
    return EINCM_OK;
}

void* eincm_iwe_fix_ptr(eincm_plan* plan) { return plan ? (void*)plan->iwe_fix : nullptr; }

int eincm_split_prepare(eincm_plan* plan, void* cuda_stream) {
    if (!plan) return EINCM_EINVAL;
    if (plan->n_peers == 0 || plan->fix_clean) return EINCM_OK;
    CU(cudaSetDevice(plan->device));
    CU(cudaMemsetAsync(plan->iwe_fix, 0, (size_t)plan->max_refs * plan->HW * sizeof(unsigned long long), (cudaStream_t)cuda_stream));
    plan->fix_clean = true;
    return EINCM_OK;
}

int eincm_split_window_images(eincm_plan* plan, void* cuda_stream) {
    if (!plan) return EINCM_EINVAL;
    if (!plan->window_set) return fail(plan, EINCM_ESTATE, "eincm_split_window_images before set_window");
    if (plan->n_peers == 0) return EINCM_OK;
    CU(cudaSetDevice(plan->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    // after the barrier: this rank's fixed-point image 0 holds the complete zero-warp image of the window
    LAUNCH("k_fix_to_f64", k_fix_to_f64<<<(int)std::min<int64_t>((plan->HW + 255) / 256, plan->sm_count * 8), 256, 0, st>>>(plan->iwe_fix, plan->HW, plan->zero_iwe));
    return EINCM_OK;
}

namespace {
static_assert(EINCM_EVAL_MAX_REFS == EINCM_MAX_REFS, "eincm_eval_metrics holds one slot per reference time");

// launches the flow-error reduction; the kFlowErrCols sums end up in dev_out (device)
int flow_error_launch(const double* pred_flow, const uint8_t* pred_mult, const double* gt_flow, const uint8_t* event_mask, int64_t n,
                      int grid, double* part, double* dev_out, cudaStream_t st) {
    k_flow_error<<<grid, 256, 0, st>>>((const double2*)pred_flow, pred_mult, (const double2*)gt_flow, event_mask, n, part);
    k_sum_rows<<<1, 32, 0, st>>>(part, grid, kFlowErrCols, dev_out);
    return cudaGetLastError() == cudaSuccess ? EINCM_OK : EINCM_ECUDA;
}

void flow_errors_from_sums(const double* s, eincm_flow_errors* out) {
    out->n_pred = (int64_t)s[0]; out->n_gt = (int64_t)s[1]; out->n_ee = (int64_t)s[2];
    out->AEE = s[2] > 0.0 ? s[3] / s[2] : std::nan("");                     // mean of an empty array
    out->AREE = s[2] > 0.0 ? s[4] / s[2] : std::nan("");
    for (int k = 0; k < 6; ++k) out->ANPE[k] = s[5 + k] * 100.0 / (s[2] + kEps);   // flow_eval.py:73-74
}
}  // namespace

int eincm_sparse_flow_error(int device, int H, int W, const double* pred_flow, const double* gt_flow, const uint8_t* event_mask,
                            eincm_flow_errors* out_host, void* cuda_stream) {
    if (!pred_flow || !gt_flow || !out_host || H < 1 || W < 1) return EINCM_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return EINCM_ECUDA;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int64_t n = (int64_t)H * W;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 1024));
    double* scratch = nullptr;
    if (cudaMalloc((void**)&scratch, ((size_t)grid + 1) * kFlowErrCols * sizeof(double)) != cudaSuccess) return EINCM_ENOMEM;
    double sums[kFlowErrCols];
    int rc = flow_error_launch(pred_flow, nullptr, gt_flow, event_mask, n, grid, scratch + kFlowErrCols, scratch, st);
    if (rc == EINCM_OK && (cudaMemcpyAsync(sums, scratch, sizeof(sums), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                           cudaStreamSynchronize(st) != cudaSuccess))
        rc = EINCM_ECUDA;
    cudaFree(scratch);
    if (rc) return rc;
    flow_errors_from_sums(sums, out_host);
    return EINCM_OK;
}

int eincm_evaluate_theta(eincm_plan* plan, const double* theta, int h, int w, const eincm_hparams* hp, const double* gt_flow,
                         const uint8_t* err_eval_event_mask, eincm_eval_metrics* out_host, void* cuda_stream) {
    if (!plan) return EINCM_EINVAL;
    if (!theta || !out_host) return fail(plan, EINCM_EINVAL, "NULL operand");
    if (plan->flags & EINCM_FLAG_EVENT_SPLIT) return fail(plan, EINCM_ESTATE, "evaluation metrics are not available on event-split plans");
    if (!plan->window_set) return fail(plan, EINCM_ESTATE, "eincm_evaluate_theta before set_window");
    int rc = check_hp(plan, hp);
    if (rc) return rc;
    CU(cudaSetDevice(plan->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    // every term is evaluated (theta_eval.py:21-42 does not gate them): force the TV and divergence paths, level 0
    eincm_hparams hpe = *hp;
    if (hpe.gamma == 0.0) hpe.gamma = 1.0;
    if (hpe.delta == 0.0) hpe.delta = 1.0;
    hpe.cur_pyr_lvl = 0;
    if ((rc = forward_events_impl(plan, theta, nullptr, 0.0, h, w, &hpe, st))) return rc;
    if ((rc = backward_impl(plan, &hpe, plan->out_stage, nullptr, nullptr, st))) return rc;
    if ((rc = ensure_theta_full(plan, st))) return rc;
    const int R = plan->R, H = plan->H, W = plan->W;
    // scratch layout inside `part`: [var of R images | var of the zero image | theta_div | flow sums | partials ...]
    double* res = plan->part;
    double* scratch = plan->part + 64;
    LAUNCH("k_image_var", k_image_var<<<R, 1024, 0, st>>>(plan->iwe, plan->HW, res, nullptr));
    LAUNCH("k_image_var(zero)", k_image_var<<<1, 1024, 0, st>>>(plan->zero_iwe, plan->HW, res + EINCM_MAX_REFS, nullptr));
    const dim3 gridD((W + kEvTX - 1) / kEvTX, (H + kEvTY - 1) / kEvTY), blockD(kEvTX, kEvTY);
    const int nD = gridD.x * gridD.y;
    if (64 + (int64_t)nD + 1 > plan->part_doubles) return fail(plan, EINCM_ENOMEM, "scratch too small for the theta divergence");
    LAUNCH("k_theta_divergence", k_theta_divergence<<<gridD, blockD, 0, st>>>((const double2*)plan->theta_full, H, W, scratch));
    LAUNCH("k_sum_rows", k_sum_rows<<<1, 32, 0, st>>>(scratch, nD, 1, res + EINCM_MAX_REFS + 1));
    if (gt_flow != nullptr) {
        const int grid = std::max(1, std::min((int)((plan->HW + 255) / 256), plan->sm_count * 4));
        if (64 + (int64_t)nD + (int64_t)grid * kFlowErrCols > plan->part_doubles) return fail(plan, EINCM_ENOMEM, "scratch too small for the flow errors");
        // pred_flow = theta_full * [pixel holds an event] (theta_eval.py:47, theta_utils.py:40-73)
        if ((rc = flow_error_launch((const double*)plan->theta_full, plan->mask, gt_flow, err_eval_event_mask, plan->HW, grid, scratch + nD,
                                    res + EINCM_MAX_REFS + 2, st)))
            return fail(plan, rc, "flow error kernels failed to launch");
        plan->launch_count += 2;
    }
    double hres[64];
    DevScalars* hs = (DevScalars*)plan->h_pinned;
    CU(cudaMemcpyAsync(hs, plan->sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(plan->h_pinned + 1024, res, 64 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::memcpy(hres, plan->h_pinned + 1024, sizeof(hres));
    eincm_eval_metrics m;
    std::memset(&m, 0, sizeof(m));
    m.n_refs = R;
    double s_con = 0.0, s_cor = 0.0, s_div = 0.0;
    for (int r = 0; r < R; ++r) {
        m.rel_contrasts[r] = hs->ref[r].contrast / (hs->zero[0].contrast + kEps);                  // losses.py:72
        m.rel_correlations[r] = (-hs->ref[r].mse) / ((-hs->zero[r].mse) + kEps);                   // losses.py:67
        m.rel_iwe_divergences[r] = hs->ref[r].div / (hs->zero[0].div + kEps);                      // losses.py:81
        m.flow_warp_losses[r] = hres[r] / hres[EINCM_MAX_REFS];                                    // losses.py:84
        m.multi_ref_weights[r] = hs->weights[r];
        s_con += m.rel_contrasts[r]; s_cor += m.rel_correlations[r]; s_div += m.rel_iwe_divergences[r];
    }
    m.mean_rel_contrast = s_con / R; m.mean_rel_corr = s_cor / R; m.mean_rel_iwe_div = s_div / R;  // theta_eval.py:27-29
    m.theta_tot_var = hs->tv_sum / (hs->tv_cnt + kEps);                                            // regularizers.py:31-36
    m.theta_div = hres[EINCM_MAX_REFS + 1] / (double)plan->HW;                                     // regularizers.py:58
    m.fwl = m.flow_warp_losses[0];                                                                 // theta_eval.py:32
    m.iwe_var = hres[0];                                                                           // theta_eval.py:36,82
    m.loss = hp->alpha * (-m.mean_rel_contrast) + hp->beta * (-m.mean_rel_corr) + hp->gamma * m.theta_tot_var + hp->delta * m.mean_rel_iwe_div;
    if (gt_flow != nullptr) {
        m.has_flow = 1;
        m.n_pixels = plan->HW;
        flow_errors_from_sums(hres + EINCM_MAX_REFS + 2, &m.flow);
    }
    *out_host = m;
    return EINCM_OK;
}

int eincm_plan_host_times(eincm_plan* plan, double* enqueue_s, double* wait_s, int64_t* n_evals, int reset) {
    if (!plan) return EINCM_EINVAL;
    if (enqueue_s) *enqueue_s = plan->host_enqueue_s;
    if (wait_s) *wait_s = plan->host_wait_s;
    if (n_evals) *n_evals = plan->host_evals;
    if (reset) { plan->host_enqueue_s = plan->host_wait_s = 0.0; plan->host_evals = 0; }
    return EINCM_OK;
}

int64_t eincm_plan_launch_count(const eincm_plan* plan) { return plan ? plan->launch_count : 0; }

int eincm_plan_set_timing(eincm_plan* plan, int enabled) {
    if (!plan) return EINCM_EINVAL;
    plan->timing = enabled != 0;
    return EINCM_OK;
}

int eincm_plan_get_timing(eincm_plan* plan, char* names_out, int names_cap, double* ms_out, int64_t* launches_out, int max_kernels,
                          int* n_kernels_out) {
    if (!plan || !n_kernels_out) return EINCM_EINVAL;
    CU(cudaSetDevice(plan->device));
    CU(cudaDeviceSynchronize());
    std::vector<std::string> names;
    std::vector<double> ms;
    std::vector<int64_t> cnt;
    for (auto& sp : plan->spans) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, sp.a, sp.b) != cudaSuccess) t = 0.f;
        size_t k = 0;
        for (; k < names.size(); ++k) if (names[k] == sp.name) break;
        if (k == names.size()) { names.push_back(sp.name); ms.push_back(0.0); cnt.push_back(0); }
        ms[k] += t; cnt[k] += 1;
        plan->event_pool.push_back(sp.a); plan->event_pool.push_back(sp.b);
    }
    plan->spans.clear();
    const int n = (int)std::min<size_t>(names.size(), (size_t)std::max(0, max_kernels));
    std::string joined;
    for (int k = 0; k < n; ++k) {
        if (ms_out) ms_out[k] = ms[k];
        if (launches_out) launches_out[k] = cnt[k];
        joined += names[k]; joined += '\n';
    }
    if (names_out && names_cap > 0) { std::snprintf(names_out, (size_t)names_cap, "%s", joined.c_str()); }
    *n_kernels_out = n;
    return EINCM_OK;
}

int eincm_debug_rounded_pixels(eincm_plan* plan, int ref, int32_t* cols_out, int32_t* rows_out, void* cuda_stream) {
    if (!plan || !cols_out || !rows_out) return EINCM_EINVAL;
    if (!plan->forward_done) return fail(plan, EINCM_ESTATE, "no evaluation yet");
    if (ref < 0 || ref >= plan->R) return fail(plan, EINCM_EINVAL, "ref %d outside 0..%d", ref, plan->R - 1);
    CU(cudaSetDevice(plan->device));
    if (plan->n_events == 0) return EINCM_OK;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    {
        const int rc = ensure_theta_full(plan, st);
        if (rc) return rc;
    }
    if (plan->exact)
        LAUNCH("k_rounded_pixels", k_rounded_pixels<true><<<event_grid(plan, plan->n_stream, 256), 256, 0, st>>>(
            plan->ev_xy, plan->ev_t, plan->perm, plan->n_stream, plan->theta_full, plan->H, plan->W, plan->tref.t[ref], cols_out, rows_out));
    else
        LAUNCH("k_rounded_pixels", k_rounded_pixels<false><<<event_grid(plan, plan->n_stream, 256), 256, 0, st>>>(
            plan->ev_xy, plan->ev_t, plan->perm, plan->n_stream, plan->theta_full, plan->H, plan->W, plan->tref.t[ref], cols_out, rows_out));
    return EINCM_OK;
}

#include "eincm_edges.inl"
#include "eincm_preproc.inl"
#include "eincm_batch.inl"
#include "eincm_ingest.inl"

}  // extern "C"
