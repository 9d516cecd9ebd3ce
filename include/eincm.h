/*
 * eincm.h - C-ABI of the B200-native EINCM contrast-correlation objective (value + gradient).
 *
 * This is the drop-in boundary for ONE hot path of robotic-vision-lab/Edge-Informed-Contrast-Maximization:
 * what `jit(value_and_grad(loss_func))` computes for the jaxopt ScipyMinimize / ScipyBoundedMinimize loops
 * (reference src/eincm/solver.py:165-183, invoked :209-216, :227-234, :325-335).  Every entry point cites
 * the reference interface it replaces.  The reference is pure Python on JAX; its "FFI" for this path is a
 * JAX custom call (`jax.ffi.ffi_call`) whose handler forwards to the functions below - see INTEGRATION.md
 * for the binding a maintainer would add on the reference side, and integration/xla_ffi/eincm_xla_ffi.cc for the handler.
 *
 * Conventions
 *   - plain C, no torch / JAX / C++ types; `cuda_stream` is a `cudaStream_t` passed as `void*`.
 *   - return 0 (EINCM_OK) on success, a negative EINCM_E* code otherwise; never throws, never aborts.
 *     `eincm_last_error(plan)` gives the message of the last failing call on that plan (NULL plan: the
 *     message of the last failing `eincm_plan_create` on this thread).
 *   - pointers are DEVICE pointers owned by the caller and borrowed for the duration of the call unless the
 *     parameter name ends in `_host`.  Outputs are caller-allocated.
 *   - calls are asynchronous and stream-ordered on `cuda_stream` (the caller synchronises) unless the
 *     function name ends in `_host` or the comment says "synchronous".
 *   - one plan = one device and one call in flight (thread-compatible, not thread-safe).
 *   - layouts are the reference's: theta / grad row-major [h][w][2] float64; edges [R][H][W] float64;
 *     events as SoA xs,ys int16 and ts float64 (reference src/experiments/e00/exp_mgr.py:379-388).
 *   - there is no CPU fallback: every compute entry point needs a CUDA device of compute capability 10.x.
 */
#ifndef EINCM_H_
#define EINCM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EINCM_ABI_VERSION 1

/* error codes */
#define EINCM_OK              0
#define EINCM_EINVAL        (-1)   /* bad argument (shape, NULL pointer, R > max_refs, theta larger than sensor ...) */
#define EINCM_ECUDA         (-2)   /* a CUDA runtime call failed; message holds cudaGetErrorString */
#define EINCM_ENOMEM        (-3)   /* allocation failed */
#define EINCM_ESTATE        (-4)   /* call order violated (e.g. value_and_grad before set_window) */
#define EINCM_ERANGE        (-5)   /* an event lies outside the sensor (the reference's loaders guarantee in-sensor events) */
#define EINCM_EUNSUPPORTED  (-6)   /* e.g. scale method other than bilinear */

/* plan flags (eincm_plan_create) */
#define EINCM_FLAG_NO_WRAP_NEGATIVE  0x1u  /* drop votes with negative row/col instead of wrapping them (JAX wraps: default) */
#define EINCM_FLAG_EVENT_SPLIT       0x2u  /* this plan holds 1/G of a window's events: split-phase calls below */

#define EINCM_FLAG_EXACT_F64         0x4u  /* nine float64 scatter-adds per event and image, float64 tap values (slower; matches the
                                              float64 reference to ~1e-13).  Default: fixed-point votes (2^-21 of the centre tap) into shared-
                                              memory windows, float64 coordinates (bit-exact pixel indices, order-independent image sums,
                                              objective within ~1e-7 relative) */

#define EINCM_FLAG_BLOCKING_SYNC     0x8u  /* the synchronous host entry points sleep on a blocking CUDA event instead of spinning in
                                              cudaStreamSynchronize: for hosts with fewer cores than driving threads (one thread per
                                              sequence, several sequences per GPU); costs some wake-up latency per call */

/* theta -> sensor-size resize method (reference configs/main.yaml:27 `scale_theta_to_sensor_size_method`) */
#define EINCM_METHOD_BILINEAR 0

typedef struct eincm_plan eincm_plan;

/* Keyword-bound hyper-parameters of the loss partial
 * (reference src/eincm/losses.py:115-122, configs/theta_loss_func/default.yaml:1-9). */
typedef struct eincm_hparams {
    double alpha;          /* weight of the contrast objective      (losses.py:115) */
    double beta;           /* weight of the correlation objective   (losses.py:116) */
    double gamma;          /* weight of the TV regulariser, applied only when cur_pyr_lvl <= 0 (losses.py:117,171) */
    double delta;          /* weight of the IWE-divergence objective (losses.py:118; shipped 0) */
    int32_t cur_pyr_lvl;   /* losses.py:119 */
    int32_t n_pyr_lvls;    /* losses.py:120 (carried for signature parity; unused by the arithmetic) */
    int32_t method;        /* EINCM_METHOD_* (losses.py:122) */
    int32_t reserved;
} eincm_hparams;

/* Layout of the float64 scalar block returned by eincm_get_scalars (indices into out_host). */
enum {
    EINCM_S_FINAL_LOSS = 0,             /* aux 'final_loss'                 (losses.py:196) */
    EINCM_S_MEAN_REL_CORR = 1,          /* aux 'mean_rel_corr'              (losses.py:198) */
    EINCM_S_MEAN_REL_CONTRAST = 2,      /* aux 'mean_rel_contrast'          (losses.py:199) */
    EINCM_S_MEAN_REL_IWE_DIV = 3,       /* aux 'mean_rel_iwe_divergence'    (losses.py:200) */
    EINCM_S_THETA_TV = 4,               /* aux 'theta_total_variation'      (losses.py:201; 0 when cur_pyr_lvl > 0) */
    EINCM_S_ZERO_CONTRAST = 5,          /* 'zero_contrast'                  (losses.py:71) */
    EINCM_S_ZERO_IWE_DIV = 6,           /* 'zero_iwe_divergence'            (losses.py:80; 0 unless delta != 0 was requested) */
    EINCM_S_DALPHA_HANDOVER = 7,        /* d loss / d alpha_handover of the last handover call */
    EINCM_S_PER_REF = 8,                /* then 5 blocks of max_refs doubles: contrasts, correlations (= -MSE),
                                           zero_correlations, iwe_divergences, multi_ref_weights */
    EINCM_S_HEADER = 8
};

/* ---- life cycle ---------------------------------------------------------------------------------------- */

/* Allocates all device state for windows of up to `max_events` events and `max_refs` reference times on a
 * sensor of H x W pixels (reference `sensor_size`, configs/dataset/dsec.yaml:1-5).  Synchronous. */
int eincm_plan_create(eincm_plan** out, int device, int H, int W, int64_t max_events, int max_refs, unsigned flags);
void eincm_plan_destroy(eincm_plan* plan);
const char* eincm_last_error(const eincm_plan* plan);
int eincm_abi_version(void);

/* ---- per window: replaces MultipleLevelEINCMSolver.set_datasample (reference src/eincm/solver.py:185-194)
 * Stages one event window: validates and bins the events by source pixel, builds the event mask
 * (theta_utils.py:66-71), the zero-warp IWE and its constants zero_contrast / zero_correlations
 * (losses.py:54-55,66,71) - everything that does not depend on theta, which the reference recomputes inside
 * every objective evaluation.  edge_ts_host: R float64 on the HOST (they become kernel constants). */
int eincm_plan_set_window(eincm_plan* plan, const int16_t* xs, const int16_t* ys, const double* ts, int64_t n_events,
                          const double* edges, const double* edge_ts_host, int n_refs, void* cuda_stream);
/* Same with the reference times in DEVICE memory, as a JAX custom call receives them: `edge_ts` is a traced operand of the
 * jitted objective (reference src/eincm/solver.py:209-216), so the XLA FFI handler (integration/xla_ffi/eincm_xla_ffi.cc)
 * only ever sees a device buffer.  One small device-to-host copy and one more synchronisation of `cuda_stream` per window. */
int eincm_plan_set_window_device_ts(eincm_plan* plan, const int16_t* xs, const int16_t* ys, const double* ts, int64_t n_events,
                                    const double* edges, const double* edge_ts, int n_refs, void* cuda_stream);

/* ---- per evaluation: replaces jit(value_and_grad(partial(loss_func, cur_pyr_lvl=l)))(theta, xs, ys, ts, edges, edge_ts)
 * (reference src/eincm/losses.py:108-205 differentiated w.r.t. argument 0; built by jaxopt from solver.py:165-173).
 * theta: [h][w][2]; loss_out: 1 float64; grad_out: [h][w][2] float64 (may be NULL: value only). */
int eincm_value_and_grad(eincm_plan* plan, const double* theta, int h, int w, const eincm_hparams* hp,
                         double* loss_out, double* grad_out, void* cuda_stream);

/* replaces jit(value_and_grad(partial(handover_loss_func, cur_pyr_lvl=l)))(alpha_handover, prev_theta, theta, ...)
 * (reference src/eincm/losses.py:208-276, differentiated w.r.t. the scalar argument 0; solver.py:175-183, :325-335).
 * dalpha_out: 1 float64 (may be NULL). */
int eincm_handover_value_and_grad(eincm_plan* plan, double alpha_handover, const double* prev_theta, const double* theta,
                                  int h, int w, const eincm_hparams* hp, double* loss_out, double* dalpha_out,
                                  void* cuda_stream);

/* ---- host-buffer forms (synchronous; H2D / D2H copies inside the call) ---------------------------------- */

/* What jaxopt's `scipy_fun(x_np) -> (value, grad)` does per line-search step (theta in, loss + grad out). */
int eincm_value_and_grad_host(eincm_plan* plan, const double* theta_host, int h, int w, const eincm_hparams* hp,
                              double* loss_out_host, double* grad_out_host, void* cuda_stream);
int eincm_handover_value_and_grad_host(eincm_plan* plan, double alpha_handover, const double* prev_theta_host,
                                       const double* theta_host, int h, int w, const eincm_hparams* hp,
                                       double* loss_out_host, double* dalpha_out_host, void* cuda_stream);
/* Batch of independent windows (BASELINE.json configs[2], [3]: "batch of windows on 1xB200", windows sharded): evaluates
 * plans[k] at thetas_host[k] ([h][w][2] each), every plan on its own internal stream, and returns when all are done.  The fixed
 * launch / synchronisation latency of a host call is paid once per batch and kernels of different windows overlap.  All plans on
 * one device; grads_out_host (or single entries of it) may be NULL: value only.  The reference evaluates windows one at a time
 * (src/eincm/solver.py:209-216); independent sequences / windows without handover may be evaluated together. */
int eincm_value_and_grad_host_batch(eincm_plan* const* plans, int n_plans, const double* const* thetas_host, int h, int w,
                                    const eincm_hparams* hp, double* losses_out_host, double* const* grads_out_host);
/* ---- batched evaluation: B staged windows, ONE launch per kernel of the evaluation (blockIdx.y = window) ----------------------
 * BASELINE.json configs[2] / [3] ("batch of windows on 1 x B200", windows sharded over the GPUs).  The reference evaluates one window
 * at a time (src/eincm/solver.py:209-216); independent windows - different sequences, or windows solved without handover - have no
 * data dependence, so their evaluations share launches here: five launches per BATCH instead of five per window, the small
 * image-space kernels fill the GPU, and MVSEC-sized windows (30 k events on 256 x 336 pixels) stop being launch-bound.  Same kernel
 * bodies as eincm_value_and_grad: results are identical bit for bit.  All plans on one device with one sensor size and the same
 * number of reference times; default path only (no EINCM_FLAG_EXACT_F64 / EVENT_SPLIT); hp->delta == 0 and no TV term (gamma == 0
 * or cur_pyr_lvl > 0); theta is a tile field of <= 4096 elements or the dense H x W field (EINCM_EUNSUPPORTED otherwise: use the
 * per-plan calls).  The pointer ARRAYS are host arrays; thetas[k] / loss_out[k] / grad_out[k] are DEVICE pointers of window k
 * ([h][w][2], 1, [h][w][2] float64).  Asynchronous on cuda_stream; one call in flight per batch. */
typedef struct eincm_batch eincm_batch;
int eincm_batch_create(eincm_batch** out, eincm_plan* const* plans, int n_plans);
void eincm_batch_destroy(eincm_batch* batch);
const char* eincm_batch_last_error(const eincm_batch* batch);
int eincm_batch_value_and_grad(eincm_batch* batch, const double* const* thetas, int h, int w, const eincm_hparams* hp,
                               double* const* loss_out, double* const* grad_out, void* cuda_stream);
/* the same with HOST operands (synchronous): one host -> device copy of all thetas, the launches, one device -> host copy of all
 * losses and gradients.  grads_out_host (or entries of it) may be NULL. */
int eincm_batch_value_and_grad_host(eincm_batch* batch, const double* const* thetas_host, int h, int w, const eincm_hparams* hp,
                                    double* losses_out_host, double* const* grads_out_host, void* cuda_stream);
int64_t eincm_batch_launch_count(const eincm_batch* batch);   /* kernels launched by this batch since creation */
/* measurement hooks: CUDA events around every launch of the batch; get_timing (synchronous) returns the milliseconds and launches per
 * kernel - splat, image statistics, image gradient, backward, theta gradient - since the last call */
int eincm_batch_set_timing(eincm_batch* batch, int enabled);
int eincm_batch_get_timing(eincm_batch* batch, double* ms_out /* [5] */, int64_t* launches_out /* [5] */);

/* ---- native optimizers: what jaxopt's ScipyMinimize(method='BFGS').run / ScipyBoundedMinimize(method='L-BFGS-B').run do with
 * the objective (reference src/eincm/solver.py:165-183, :209-216, :325-335), without returning to Python between evaluations.
 * scipy.optimize.minimize restated (csrc/eincm_opt.h): BFGS with scipy's line-search policy (MINPACK-2 dcsrch, then scalar_search_wolfe2
 * when that fails) and the n = 1 case of L-BFGS-B 3.0 (dcsrch with xtol 0.1, restart from steepest descent after a failed search);
 * tests/test_native_opt.py holds them to scipy's iterations, evaluations and final points.  status as scipy reports it: 0 converged
 * (max|grad| <= gtol), 1 maxiter reached, 2 "precision loss" (both line searches failed or the objective is not finite; L-BFGS-B:
 * abnormal termination in the line search), 3 NaN in the result - the codes solver.py:218-239 reacts to.
 * eincm_minimize_bfgs_host keeps a dense (2hw)^2 inverse Hessian on the host: h * w <= 4096, EINCM_EINVAL beyond.
 * cuda_stream == (void*)-1 selects the plan's own stream (several plans solved concurrently from several host threads). */
typedef struct eincm_opt_result { double fun; int32_t nit, nfev, status, reserved; } eincm_opt_result;
int eincm_minimize_bfgs_host(eincm_plan* plan, double* theta_inout_host /* [h][w][2] */, int h, int w, const eincm_hparams* hp,
                             int maxiter, double gtol, eincm_opt_result* result_out, void* cuda_stream);
int eincm_minimize_handover_host(eincm_plan* plan, double* alpha_inout_host, double lo, double hi, const double* prev_theta_host,
                                 const double* theta_host, int h, int w, const eincm_hparams* hp, int maxiter, double pgtol,
                                 eincm_opt_result* result_out, void* cuda_stream);
/* The same BFGS level solve with the loop ON THE DEVICE.  k_bfgs_step consumes the loss and gradient of the evaluation
 * that just ran at a trial point in device memory, advances scipy's line search (the same resumable machines as eincm_minimize_bfgs_host:
 * csrc/eincm_linesearch.h), applies the inverse-Hessian update and writes the next trial point.  Two CUDA-graph forms: unrolled (default) -
 * 8 x { evaluation kernels ; k_bfgs_step } as an ordinary graph that the call relaunches until the level is done (~7 launches per level;
 * the kernels behind the end of the level return at once), graphs of several plans overlap like streams; or one WHILE conditional node
 * around { evaluation ; k_bfgs_step } (environment EINCM_GRAPH_UNROLL=0, and whenever the TV regulariser is active): one launch per level.
 * Same arguments, results and status codes as eincm_minimize_bfgs_host (iterates agree to rounding: the reductions run in another order);
 * h * w * 2 <= 1024 flow parameters, default evaluation path only (no EINCM_FLAG_EXACT_F64, delta == 0), EINCM_EINVAL / EINCM_ESTATE
 * otherwise.  cuda_stream: NULL or (void*)-1 select the plan's own stream (the legacy default stream cannot be captured).  With
 * EINCM_FLAG_BLOCKING_SYNC the host sleeps between launches: no spinning core per sequence.  The graph of a level is rebuilt once per
 * staged window (it bakes the window's reference times and chunk count) and re-used for repeated solves of that window. */
int eincm_minimize_bfgs_graph_host(eincm_plan* plan, double* theta_inout_host /* [h][w][2] */, int h, int w, const eincm_hparams* hp,
                                   int maxiter, double gtol, eincm_opt_result* result_out, void* cuda_stream);
/* The BATCHED form of the device-side loop ("batched BFGS driver for many windows"; reference src/eincm/solver.py:165-173,
 * 209-216 runs one scipy BFGS per window): all windows of an eincm_batch solve one pyramid level in lockstep - an unrolled graph of
 * 8 x { the five batched evaluation kernels (blockIdx.y = window) ; k_bfgs_step_b (one CTA per window) } relaunched until no window is left.
 * Every window follows the schedule eincm_minimize_bfgs_graph_host would run for it alone; a window whose level has ended is skipped by
 * the remaining kernels (they read its `done` word), so the GPU time of a step is proportional to the windows still running.  The graph
 * depends on (theta shape, R, grid size) only - the windows' operands live in the batch's argument records - and is re-used across
 * batches of windows.  thetas_inout_host: [n_plans][h][w][2] contiguous; results_out: [n_plans]; active: NULL, or [n_plans] flags -
 * windows with a zero flag take no part (their theta and result are left untouched: the retries of solver.py:218-226 re-solve a subset).
 * Same limits as the single-window loop and as eincm_batch_value_and_grad (gamma == 0 or cur_pyr_lvl > 0, delta == 0).  Synchronous. */
int eincm_batch_minimize_bfgs_graph_host(eincm_batch* batch, double* thetas_inout_host, int h, int w, const eincm_hparams* hp, int maxiter,
                                         double gtol, const int32_t* active, eincm_opt_result* results_out, void* cuda_stream);
int64_t eincm_batch_solve_launches(const eincm_batch* batch);   /* graph launches of the last batched solve */
/* Stateless single shot with the exact operand list of loss_func (losses.py:108-114), every operand on the host:
 * set_window + value_and_grad + copies. */
int eincm_value_and_grad_stateless_host(eincm_plan* plan, const double* theta_host, int h, int w,
                                        const int16_t* xs_host, const int16_t* ys_host, const double* ts_host, int64_t n_events,
                                        const double* edges_host, const double* edge_ts_host, int n_refs,
                                        const eincm_hparams* hp, double* loss_out_host, double* grad_out_host,
                                        void* cuda_stream);

/* ---- split-phase form for one window whose events are split over G GPUs (EINCM_FLAG_EVENT_SPLIT) --------
 * The splat is additive in events; everything after it needs the complete image.  Sequence per rank:
 *   set_window (local events)      -> all-reduce(sum) eincm_zero_iwe_ptr  [H*W]   -> eincm_window_finalize
 *                                     (+ all-reduce(max) eincm_mask_ptr [H*W] when gamma != 0)
 *   eincm_forward_events (local)   -> all-reduce(sum) eincm_iwe_ptr       [R*H*W] -> eincm_backward
 *   -> all-reduce(sum) of grad_out (and of the gamma-free loss is identical on all ranks already)
 * The collective itself is the caller's (torch.distributed / NCCL over NVLink); these functions only expose
 * the buffers.  Without the flag, set_window / value_and_grad run all phases back to back. */
int eincm_window_finalize(eincm_plan* plan, void* cuda_stream);
/* Position of this plan in the split: the gradient of terms that are replicated on every rank (the TV regulariser, which
 * depends on theta and the all-reduced mask only) is added by rank 0 alone, so that all-reduce(sum) of grad_out is exact. */
int eincm_plan_set_event_split(eincm_plan* plan, int rank, int world);
int eincm_forward_events(eincm_plan* plan, const double* theta, int h, int w, const eincm_hparams* hp, void* cuda_stream);
int eincm_backward(eincm_plan* plan, const eincm_hparams* hp, double* loss_out, double* grad_out, void* cuda_stream);
/* Fixed-point form of the per-evaluation collective (on != 0): eincm_forward_events leaves this rank's votes in the FIXED-POINT images
 * and the caller all-reduces(sum, int64) eincm_iwe_fix_ptr [R*H*W] instead of the float64 images.  Integer sums do not depend on the
 * order of the reduction: every rank holds bit-identical complete images whatever algorithm the collective uses, eincm_backward runs
 * the fused image pass on them directly (no float64 copy, two image kernels instead of five), and the objective is bit-identical on
 * all ranks and equal to the single-GPU value.  (delta != 0 falls back to the float64 images.) */
int eincm_plan_set_split_fixed_point(eincm_plan* plan, int on);
/* ---- event split with peer access: the all-reduce of the partial images fused into the splat --------------------
 * One process per GPU of one NVLink / NVSwitch domain.  Every rank publishes the CUDA IPC handle of its fixed-point image
 * buffer (eincm_plan_ipc_handle, 64 bytes), the caller exchanges the handles (torch.distributed all-gather) and hands all of
 * them, in rank order, to eincm_plan_set_peers (after eincm_plan_set_event_split).  From then on the splat adds the non-zero
 * cells of its shared-memory windows to the image of EVERY rank with 64-bit integer reductions over NVLink: integer sums are
 * order-independent, so after a cross-rank barrier every rank holds the identical complete images and runs the (fused) image
 * pass on them - no all-reduce of the R*H*W images, and the objective is bit-identical on all ranks.  Sequence per rank:
 *   eincm_split_prepare  -> barrier -> set_window (local events) -> barrier -> eincm_split_window_images -> eincm_window_finalize
 *   eincm_split_prepare  -> barrier -> eincm_forward_events      -> barrier -> eincm_backward -> all-reduce(sum) of grad_out
 * (eincm_split_prepare clears this rank's image buffer when it is not clean already; the barrier before the splat guarantees
 * that no rank adds into a buffer that is still being read or cleared - the gradient all-reduce of the previous evaluation can
 * serve as that barrier.)  eincm_plan_set_peer_pointers is the in-process form (plans of one process, e.g. tests). */
int eincm_plan_ipc_handle(eincm_plan* plan, void* handle_out, int handle_bytes);
int eincm_plan_set_peers(eincm_plan* plan, const void* handles /* [world][64] */, int n_handles);
int eincm_plan_set_peer_pointers(eincm_plan* plan, void* const* fix_ptrs /* [world], eincm_iwe_fix_ptr of every rank */, int n_ptrs);
void* eincm_iwe_fix_ptr(eincm_plan* plan);       /* device, [max_refs*H*W] uint64 fixed-point images (2^21 * 2 pi * value) */
int eincm_split_prepare(eincm_plan* plan, void* cuda_stream);
int eincm_split_window_images(eincm_plan* plan, void* cuda_stream);
double* eincm_zero_iwe_ptr(eincm_plan* plan);   /* device, [H*W] float64 */
double* eincm_iwe_ptr(eincm_plan* plan);        /* device, [R*H*W] float64, image of warped events of the last evaluation */
/* device, [H*W] uint8: 1 where a pixel holds >= 1 (local) event (theta_utils.py:66-71); event-split plans that use
 * gamma != 0 all-reduce(max) it together with the zero-IWE */
uint8_t* eincm_mask_ptr(eincm_plan* plan);

/* ---- read-outs (debug taps and the aux dict of loss_func) ------------------------------------------------ */
double* eincm_dldi_ptr(eincm_plan* plan);        /* device, [R*H*W]: d loss / d IWE of the last evaluation */
double* eincm_theta_full_ptr(eincm_plan* plan);  /* device, [H*W*2]: aux 'scaled_theta' (losses.py:197) */
/* synchronous: copies EINCM_S_HEADER + 5*max_refs doubles (see enum above) to the host */
int eincm_get_scalars(eincm_plan* plan, double* out_host, int n_doubles, void* cuda_stream);
/* Bit-exact event->pixel index stream of reference r for the last evaluated theta, in the ORIGINAL event
 * order: cols_out/rows_out [n_events] int32 = Xs_rounded of event_utils.py:33 (before the +dx,+dy taps). */
int eincm_debug_rounded_pixels(eincm_plan* plan, int ref, int32_t* cols_out, int32_t* rows_out, void* cuda_stream);
/* ---- evaluation groups: several sequences solved concurrently on one GPU, launched by one thread ---------------------------
 * The reference solves one window at a time; on a B200 several independent sequences (one plan, one host thread each) are solved
 * concurrently so that kernels of different windows overlap.  Plans that joined the same group rendezvous inside
 * eincm_minimize_bfgs_host: a thread that needs an objective evaluation posts it and sleeps; when every member that is currently
 * inside a minimisation has posted, one of them launches all the posted evaluations back to back (each on its plan's own stream),
 * collects them and wakes the others.  Launches then come in bursts from ONE thread (the evaluations of a burst overlap on the GPU
 * like eincm_value_and_grad_host_batch) and the waiting threads do not spin - the form to use when there are fewer host cores
 * than sequences.  Results are identical to the ungrouped calls (each optimizer sees exactly its own evaluations). */
typedef struct eincm_group eincm_group;
int eincm_group_create(eincm_group** out);
void eincm_group_destroy(eincm_group* group);
/* A burst is launched as soon as `percent` % of the members that are inside a minimisation have posted (default 100: lock step).
 * With 50 and twice as many sequences, one half evaluates on the GPU while the optimizers of the other half do their host-side
 * work (line search, BFGS update). */
int eincm_group_set_burst_percent(eincm_group* group, int percent);
/* group == NULL leaves the group.  All plans of a group must live on one device.  Not thread-safe against a running minimisation
 * of the same plan. */
int eincm_plan_set_group(eincm_plan* plan, eincm_group* group);

/* ---- evaluation metrics of a solved window (reference src/evaluations/theta_eval.py:14-95, flow_eval.py:14-75) ----------
 * Computed on the device from the images the evaluation leaves there; one call per solved window (not the optimisation hot path). */
#define EINCM_EVAL_MAX_REFS 8
typedef struct {
    double AEE, AREE;            /* mean end-point error / relative end-point error over the valid pixels (NaN when n_ee == 0) */
    double ANPE[6];              /* percentage of valid pixels with end-point error > 1, 2, 3, 5, 10, 20 px (flow_eval.py:72-74) */
    int64_t n_ee, n_pred, n_gt;  /* counts of flow_eval.py:65-67 */
} eincm_flow_errors;

typedef struct {
    double loss;                 /* alpha (-mean_rel_contrast) + beta (-mean_rel_corr) + gamma tot_var + delta mean_rel_iwe_div (theta_eval.py:37-42) */
    double iwe_var;              /* var of the image of warped events at the first reference time (theta_eval.py:36,82) */
    double mean_rel_contrast, mean_rel_corr, mean_rel_iwe_div;   /* UN-weighted means over the reference times (theta_eval.py:27-29) */
    double theta_tot_var, theta_div;                             /* regularizers.py:14-38 / :41-58 */
    double fwl;                  /* flow warp loss at the first reference time: var(IWE_0) / var(zero-IWE) (contrast_metrics.py:6-17) */
    double rel_contrasts[EINCM_EVAL_MAX_REFS], rel_correlations[EINCM_EVAL_MAX_REFS], rel_iwe_divergences[EINCM_EVAL_MAX_REFS];
    double flow_warp_losses[EINCM_EVAL_MAX_REFS], multi_ref_weights[EINCM_EVAL_MAX_REFS];
    int32_t n_refs, has_flow;    /* has_flow != 0: `flow` and n_pixels are valid (a ground-truth flow was given) */
    int64_t n_pixels;
    eincm_flow_errors flow;
} eincm_eval_metrics;

/* sparse_flow_error(pred_flow, gt_flow, event_mask) of flow_eval.py:14-75.  pred_flow / gt_flow: DEVICE [H*W*2] float64 (row-major
 * [H][W][2]); event_mask: DEVICE [H*W] uint8 or NULL.  Synchronous; result in host memory. */
int eincm_sparse_flow_error(int device, int H, int W, const double* pred_flow, const double* gt_flow, const uint8_t* event_mask,
                            eincm_flow_errors* out_host, void* cuda_stream);
/* evaluate_theta_array of theta_eval.py:14-95 on the staged window.  theta: DEVICE [h][w][2] (up-scaled to the sensor size like
 * loss_func does; pass h = H, w = W for a dense field).  gt_flow: DEVICE [H][W][2] or NULL; err_eval_event_mask: DEVICE [H*W] uint8
 * or NULL.  The total-variation and divergence terms are always evaluated (the reference's evaluation does not gate them on
 * cur_pyr_lvl, gamma or delta); hp supplies alpha, beta, gamma, delta.  Synchronous.  Not for event-split plans. */
int eincm_evaluate_theta(eincm_plan* plan, const double* theta, int h, int w, const eincm_hparams* hp, const double* gt_flow,
                         const uint8_t* err_eval_event_mask, eincm_eval_metrics* out_host, void* cuda_stream);

/* ---- edge images of a window from its grayscale frames (the step before set_window) --------
 * Replaces, per reference time, `normalize_to_unit_range(smoothen_edges(image_to_edge(jnp_to_ocv_n255(image))))` of
 * src/experiments/e00/exp_mgr.py:343-350:
 *   image_to_edge                = cv.Canny(img, th1, th2, None, 3, L2gradient=True)            src/utils/img_utils.py:194-211
 *   smoothen_edges               = cv.GaussianBlur(float64 edge image, None, k_size, sigma, 0)  src/utils/img_utils.py:213-222
 *                                  (OpenCV reads k_size as sigmaX and derives the kernel size from it; `sigma` is ignored)
 *   eincm_inv_exp_dist_transform = 1 - normalize(1 - exp(-EDT(~edge) / alpha))                  src/utils/img_utils.py:231-235
 * The Canny stage is bit-exact with OpenCV (integer arithmetic); the smoothing stages are float64 (1e-12).  The image
 * pre-processing before Canny (non-local-means denoise, CLAHE, sharpen, bilateral filter: preprocess_image, img_utils.py:131-191)
 * stays with OpenCV: `images` are the uint8 frames that cv.Canny receives. */
#define EINCM_SMOOTHEN_GAUSSIAN 0      /* configs/edge_extraction/smoothen/gaussian.yaml */
#define EINCM_SMOOTHEN_IEDT     1      /* configs/edge_extraction/smoothen/iedt.yaml */
#define EINCM_EDGE_STAGE_CANNY      1
#define EINCM_EDGE_STAGE_SMOOTHEN   2  /* always performed */
#define EINCM_EDGE_STAGE_NORMALIZE  4
typedef struct eincm_edge_params {
    double canny_th1, canny_th2;       /* edge_extraction.canny.threshold_1 / threshold_2 (run.sh: 30 / 80 DSEC, 100 / 200 MVSEC) */
    int32_t smoothen;                  /* EINCM_SMOOTHEN_* */
    int32_t stages;                    /* 0 = all; else a mask of EINCM_EDGE_STAGE_*: without CANNY `images` are taken as edge images
                                          (the stand-alone smoothen_edges / eincm_inv_exp_dist_transform), without NORMALIZE the
                                          smoothing function's own result is returned */
    double gauss_sigma;                /* smoothen_edges' k_size (= sigmaX as OpenCV reads the call); gaussian.yaml: 1 */
    double iedt_alpha;                 /* eincm_inv_exp_dist_transform's alpha; iedt.yaml: 6 / 5.541 */
} eincm_edge_params;
/* bytes of device scratch eincm_edge_maps needs for n_images frames of H x W (0 for invalid sizes) */
size_t eincm_edge_workspace_bytes(int H, int W, int n_images);
/* images: DEVICE [n_images][H][W] uint8; edges_out: DEVICE [n_images][H][W] float64 (the `edges` operand of
 * eincm_plan_set_window); canny_out: DEVICE [n_images][H][W] uint8 or NULL (cv.Canny's 0 / 255 image, debug tap);
 * workspace: DEVICE, eincm_edge_workspace_bytes.  Asynchronous on cuda_stream. */
int eincm_edge_maps(int device, const uint8_t* images, int n_images, int H, int W, const eincm_edge_params* p, double* edges_out,
                    uint8_t* canny_out, void* workspace, size_t workspace_bytes, void* cuda_stream);
/* the same with HOST buffers (allocates, copies, synchronises): what a loader thread calls once per window */
int eincm_edge_maps_host(int device, const uint8_t* images_host, int n_images, int H, int W, const eincm_edge_params* p,
                         double* edges_out_host, uint8_t* canny_out_host);

/* Non-local-means denoise of the frames, the first and by far the slowest step of preprocess_image (src/utils/img_utils.py:147-157:
 * cv.fastNlMeansDenoising(img, None, h, template_win_size, search_win_size); denoise/default.yaml: h 4, template 3, search 11 -
 * ~110 ms per 640x480 frame with OpenCV on 8 host threads).  uint8, one channel, bit-exact with OpenCV (integer patch distances,
 * OpenCV's fixed-point weight table, BORDER_REFLECT_101).  images / out: DEVICE [n_images][H][W] uint8 (must not alias);
 * workspace: DEVICE, eincm_nlm_workspace_bytes (weight table).  Window sizes are made odd like OpenCV does (size / 2 * 2 + 1);
 * template <= 15, search <= 41.  Asynchronous on cuda_stream. */
size_t eincm_nlm_workspace_bytes(int template_window_size, int search_window_size);
int eincm_nlm_denoise(int device, const uint8_t* images, int n_images, int H, int W, float h, int template_window_size,
                      int search_window_size, uint8_t* out, void* workspace, size_t workspace_bytes, void* cuda_stream);

/* ---- image pre-processing between the denoise and the bilateral filter (src/utils/img_utils.py:159-181) --------
 * uint8 frames, DEVICE [n_images][H][W], bit-exact with OpenCV 4.x (restated and pinned against cv2 in oracle/edge_oracle.py).
 * eincm_clahe: cv.createCLAHE(clipLimit, tileGridSize = (tiles_x, tiles_y)).apply (img_utils.py:159-161) - per-tile histogram, clip +
 * redistribution, LUT, float32 bilinear blend of the four neighbouring LUTs; workspace: DEVICE, eincm_clahe_workspace_bytes (the LUTs).
 * eincm_sharpen: cv.GaussianBlur of the uint8 frame (8-bit fixed-point kernel of size round(6 sigma + 1) | 1 <= 63, BORDER_REFLECT_101)
 * followed by cv.addWeighted(frame, alpha, blur, beta, gamma) in float32 (img_utils.py:163-178; as OpenCV reads the reference's positional
 * call, sigma is `sharpen_kernel_size`); blur_out: optional copy of the blurred frame; out must not alias images.
 * The bilateral filter that follows (img_utils.py:183-189) stays an OpenCV call: in the opencv-python build its interior comes from Intel
 * IPP and its border columns from OpenCV's own code, so its output is a property of the build.  Asynchronous on cuda_stream. */
size_t eincm_clahe_workspace_bytes(int n_images, int tiles_x, int tiles_y);
int eincm_clahe(int device, const uint8_t* images, int n_images, int H, int W, double clip_limit, int tiles_x, int tiles_y, uint8_t* out,
                void* workspace, size_t workspace_bytes, void* cuda_stream);
int eincm_sharpen(int device, const uint8_t* images, int n_images, int H, int W, double sigma, double alpha, double beta, double gamma,
                  uint8_t* blur_out, uint8_t* out, void* cuda_stream);

/* ---- event ingest before staging --------
 * What the reference's DSEC loader / experiment manager do in NumPy between the h5 event stream and loss_func's operands.
 * All pointers DEVICE unless suffixed _host. */
/* rectify_events, src/dataloaders/dsec_loader.py:145-170: (x, y) <- round-half-even(rectify_map[y, x]) as int16; events whose
 * rectified pixel leaves the H x W sensor are dropped, the order of the others is kept.  rectify_map: float32 [H][W][2] (x, y).
 * t / t_out (int64 microseconds) and p / p_out (uint8 polarity) may be NULL together.  Outputs hold up to n_events entries.
 * Synchronous (the surviving count is returned in host memory). */
size_t eincm_rectify_workspace_bytes(int64_t n_events);
int eincm_rectify_events(int device, const int16_t* x, const int16_t* y, const int64_t* t, const uint8_t* p, int64_t n_events,
                         const float* rectify_map, int H, int W, int16_t* x_out, int16_t* y_out, int64_t* t_out, uint8_t* p_out,
                         int64_t* n_out_host, void* workspace, size_t workspace_bytes, void* cuda_stream);
/* stage_datasample, src/experiments/e00/exp_mgr.py:313-321: ts = (t - start) / (end - start + eps) in float64 (the `ts` operand
 * of loss_func / eincm_plan_set_window).  Asynchronous. */
int eincm_normalize_times(int device, const int64_t* t_us, int64_t n_events, int64_t start_us, int64_t end_us, double* ts_out,
                          void* cuda_stream);
/* fixed-N window rule of DSECDataLoader.get_sample, dsec_loader.py:293-311 (host arithmetic): the event index range
 * [idx_start, idx_end) of a window is widened symmetrically (ceil / floor of half the deficiency, clamped to the stream) or cut to
 * des_n_events (keeping the latest or the earliest events).  des_n_events <= 0: unchanged. */
int eincm_window_event_range(int64_t idx_start, int64_t idx_end, int64_t n_total, int64_t des_n_events, int prefer_latest_events,
                             int64_t* start_out, int64_t* end_out, int64_t* deficiency_out);

/* ---- measurement hooks (bench.py): launch accounting and optional per-kernel CUDA-event timing -------- */
/* wall time the synchronous host entry points of this plan spent launching (enqueue_s) and waiting for results (wait_s) over
 * n_evals evaluations since the last reset: shows whether a solve loop is bound by the host or by the GPU */
int eincm_plan_host_times(eincm_plan* plan, double* enqueue_s, double* wait_s, int64_t* n_evals, int reset);
/* number of kernels this plan has launched since creation (memsets / copies not counted) */
int64_t eincm_plan_launch_count(const eincm_plan* plan);
/* enabled != 0: bracket every subsequent kernel launch with CUDA events on the launching stream */
int eincm_plan_set_timing(eincm_plan* plan, int enabled);
/* synchronous: sums the recorded spans per kernel name since the last call and clears them.
 * names_out: '\n'-separated kernel names (names_cap bytes); ms_out / launches_out: [max_kernels]. */
int eincm_plan_get_timing(eincm_plan* plan, char* names_out, int names_cap, double* ms_out, int64_t* launches_out, int max_kernels,
                          int* n_kernels_out);
int eincm_plan_info(const eincm_plan* plan, int* H, int* W, int64_t* n_events, int* n_refs, int* max_refs);

#ifdef __cplusplus
}
#endif
#endif /* EINCM_H_ */
