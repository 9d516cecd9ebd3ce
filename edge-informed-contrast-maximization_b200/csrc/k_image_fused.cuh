// Fused image-space pass of the default (tile-privatised splat, delta == 0) path: two kernels per evaluation
// instead of compose + A + scalars + B + scalars + C.
//
//   k_image_stats: fixed-point image -> float64 image of warped events, Scharr contrast, min / max with
//                 tie counts, and the moments  sum I, sum I^2, sum E*I  from which the min-max-normalised MSE and the sums of
//                 its backward follow algebraically once the GLOBAL min / max are known:
//                     N = (I - m)/D,  D = max - m + eps
//                     sum (E-N)^2     = sum E^2 - 2/D (sum EI - m sum E) + Q/D^2,      Q = sum I^2 - 2 m sum I + HW m^2
//                     s1 = sum gN     = cb (sum E - (sum I - HW m)/D)                   gN = cb (E - N)
//                     s2 = sum gN(I-m)= cb ((sum EI - m sum E) - Q/D)
//                 so no second pass over the images is needed for the statistics (reference: src/utils/img_utils.py:24-25,
//                 src/eincm/objectives/correlation_objectives.py:25-26, contrast_objectives.py:22-25, src/eincm/losses.py:171-193);
//                 It also applies the adjoint of the Scharr pair to (Gx, Gy) - the part of d loss / d IWE that does not depend on
//                 the global statistics - and its last CTA reduces the per-CTA partials and evaluates the loss.
//   k_image_grad:  pointwise  d loss / d IWE = cA * adjoint + gN / D + min/max tie terms  (float64 + float32/2pi copies, as
//                 k_img_C), and clears the fixed-point images for the next evaluation.
#pragma once

#include "common.cuh"
#include "k_events_tile.cuh"
#include "k_image.cuh"

namespace eincm {

constexpr int kFTX = 32, kFTY = 16, kFNT = 256;
constexpr int kFPart = 8;    // doubles per CTA partial: sq, sI, sI2, sEI, mn, cnt_mn, mx, cnt_mx

__device__ __forceinline__ void merge_min(double& v, double& c, double ov, double oc) {
    if (ov < v) { v = ov; c = oc; } else if (ov == v) { c += oc; }
}
__device__ __forceinline__ void merge_max(double& v, double& c, double ov, double oc) {
    if (ov > v) { v = ov; c = oc; } else if (ov == v) { c += oc; }
}

struct FusedAcc {
    double sq, sI, sI2, sEI, mn, cmn, mx, cmx;
    __device__ __forceinline__ void init() { sq = sI = sI2 = sEI = 0.0; mn = INFINITY; mx = -INFINITY; cmn = cmx = 0.0; }
    __device__ __forceinline__ void merge(const FusedAcc& o) {
        sq += o.sq; sI += o.sI; sI2 += o.sI2; sEI += o.sEI;
        merge_min(mn, cmn, o.mn, o.cmn);
        merge_max(mx, cmx, o.mx, o.cmx);
    }
    __device__ __forceinline__ FusedAcc shfl_xor(int o) const {
        FusedAcc r;
        r.sq = __shfl_xor_sync(0xffffffffu, sq, o); r.sI = __shfl_xor_sync(0xffffffffu, sI, o);
        r.sI2 = __shfl_xor_sync(0xffffffffu, sI2, o); r.sEI = __shfl_xor_sync(0xffffffffu, sEI, o);
        r.mn = __shfl_xor_sync(0xffffffffu, mn, o); r.cmn = __shfl_xor_sync(0xffffffffu, cmn, o);
        r.mx = __shfl_xor_sync(0xffffffffu, mx, o); r.cmx = __shfl_xor_sync(0xffffffffu, cmx, o);
        return r;
    }
};

// ---- k_image_stats ---------------------------------------------------------------------------------------------------------
// Streaming formulation, no shared memory and no block barrier in the main loop: one WARP owns a strip of kS2Cols columns x kS2Rows
// rows of one reference image and marches down its rows.  Lanes are adjacent columns (two halo lanes on either side: the Scharr pair
// and its adjoint are two chained 3x3 stencils), horizontal neighbours come from warp shuffles, vertical neighbours from a
// three-row window in registers; rows are loaded kS2Pre at a time, one group ahead of the arithmetic (all loads of a group in flight
// together).  Every warp accumulates the statistics of its own pixels and writes one partial record; the last CTA to finish
// (ticket) reduces the records of every reference image in a fixed order (deterministic) and evaluates the loss.  The kernel
// boundary before k_image_grad replaces a grid-wide barrier.  Every CTA writes ONE partial record (its warps merged in warp order); the
// last CTA to finish reduces the records, one warp per reference image, and publishes only what k_image_grad needs (stats_tail); the
// loss is evaluated by a spare CTA of k_image_grad.  History (640x480, R = 3): whole row bands in shared memory, two CTAs per SM,
// barriers between phases: 27 us; streaming warps, one record per WARP and a serial tail that also evaluated the loss: 45 us, more
// than 20 of them in the tail (ncu: 14 % of the warp slots active); the tail replicated in the prologue of every CTA of k_image_grad:
// 21 + 25 us (and slower with four windows in flight: 300 CTAs repeat it); this form: 29 us, k_image_grad 10.8 us.
constexpr int kCoopMaxRefs = EINCM_MAX_REFS;
constexpr int kS2Warps = 4;                         // warps (work items) per CTA
constexpr int kS2NT = kS2Warps * 32;
constexpr int kS2Cols = 28;                         // own columns of a warp (lanes 2 .. 29)
#ifndef EINCM_S2_ROWS
#define EINCM_S2_ROWS 12
#endif
constexpr int kS2Rows = EINCM_S2_ROWS;                         // own rows of a warp (640x480, R = 3: 12 rows 140.8 us per evaluation on one stream, 24 rows 144.9; kernel 29 us either way)
#ifndef EINCM_S2_MINB
#define EINCM_S2_MINB 5
#endif
constexpr int kS2Pre = 4;                           // rows per load group
constexpr int kStatsTicket = 6;                     // DevScalars::counters slot of the "last CTA" ticket

struct ImageStatsArgs {
    const unsigned long long* fix;  // [R][H*W] fixed-point images of warped events
    const double* edges;            // [R][H*W]
    double* iwe;                    // [R][H*W] float64 images (out)
    float* adj32;                   // [R][H*W] adjoint of the Scharr pair applied to (Gx, Gy) (out; d contrast / d IWE up to cA)
    double* part;                   // [R][image_stats_ctas(H, W)][kFPart]: one record per CTA
    DevScalars* sc;
    double* loss_out;
    const int* skip;                // non-zero: return at once (unrolled solve graphs), or null
    double* zero_buf;               // buffers cleared for the event backward pass (dense flow-field gradient, theta gradient) or null
    double* zero_buf2;
    int n_zero, n_zero2;
    int H, W, R;
    double alpha, beta, gamma;
    int use_tv;
    int pad;
};

__host__ __device__ inline int image_stats_strips(int W) { return (W + kS2Cols - 1) / kS2Cols; }
__host__ __device__ inline int image_stats_bands(int H) { return (H + kS2Rows - 1) / kS2Rows; }
// work items (warps) of the image pass for R reference images
__host__ __device__ inline int image_stats_items(int H, int W, int R) { return R * image_stats_strips(W) * image_stats_bands(H); }
// CTAs per reference image: kS2Warps work items each, never items of two images (one partial record per CTA)
__host__ __device__ inline int image_stats_ctas(int H, int W) { return (image_stats_strips(W) * image_stats_bands(H) + kS2Warps - 1) / kS2Warps; }

__device__ __forceinline__ Stats fused_stats(const FusedAcc& a, double n, double sE, double sE2, double cb) {
    Stats st;
    const double m = a.mn, D = (a.mx - a.mn) + kEps;                      // img_utils.py:25
    const double EIm = a.sEI - m * sE;                                   // sum E (I - m)
    const double Q = a.sI2 - 2.0 * m * a.sI + n * m * m;                 // sum (I - m)^2
    st.contrast = a.sq / n;
    st.mn = m; st.mx = a.mx; st.D = D;
    st.mse = (sE2 - 2.0 * EIm / D + Q / (D * D)) / n;
    st.s1 = cb * (sE - (a.sI - n * m) / D);
    st.s2 = cb * (EIm - Q / D);
    st.cnt_min = a.cmn; st.cnt_max = a.cmx;
    st.div = 0.0;
    return st;
}

struct StatsTail {
    double wacc[kS2Warps][kFPart];     // per-warp partials of one CTA of k_image_stats
    int is_last;
};

// Tail of the image statistics, run by the last CTA of k_image_stats to finish: reduces the per-CTA records of every reference image
// - one WARP per image, records in lane order, butterfly merge (commutative operations: every lane holds the same bits, the result
// does not depend on which CTA runs it) - and publishes the per-image statistics and cotangent scales k_image_grad needs.  Nothing
// else: the loss is a serial chain of float64 divisions that nothing before k_theta_grad waits for (publish_loss, run by a spare CTA of
// k_image_grad next to the pointwise pass).
template <int NT>
__device__ __forceinline__ void stats_tail(const double* __restrict__ part, DevScalars* sc, int R, int cpi, int HW, double alpha, double beta) {
    const int tid = linear_tid(), lane = tid & 31, wid = tid >> 5;
    for (int q = wid; q < R; q += NT / 32) {
        // per-window constants of the image (cotangent scales from the zero-warp image, losses.py:176-177): loads in flight with the records
        const double w = sc->weights[q], zc = sc->zero[0].contrast, zm = sc->zero[q].mse, sE = sc->sumE[q], sE2 = sc->sumE2[q];
        FusedAcc a;
        a.init();
        for (int k0 = lane; k0 < cpi; k0 += 4 * 32) {
            double v[4][kFPart];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double* d = part + ((int64_t)q * cpi + min(k0 + 32 * u, cpi - 1)) * kFPart;
#pragma unroll
                for (int f = 0; f < kFPart; ++f) v[u][f] = __ldcg(d + f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (k0 + 32 * u < cpi) {
                    FusedAcc o;
                    o.sq = v[u][0]; o.sI = v[u][1]; o.sI2 = v[u][2]; o.sEI = v[u][3];
                    o.mn = v[u][4]; o.cmn = v[u][5]; o.mx = v[u][6]; o.cmx = v[u][7];
                    a.merge(o);
                }
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) a.merge(a.shfl_xor(o));
        if (lane == 0) {
            const double a_r = -alpha * w / ((zc + kEps) * R);
            const double b_r = beta * w / ((-zm + kEps) * R);
            const double cA = a_r * (2.0 / (double)HW), cB = b_r * (-2.0 / (double)HW);
            sc->ref[q] = fused_stats(a, (double)HW, sE, sE2, cB);
            sc->coefA[q] = cA; sc->coefB[q] = cB; sc->coefD[q] = 0.0;
        }
    }
}

// Loss from the published statistics (one thread; reference src/eincm/losses.py:171-193).
__device__ __forceinline__ void publish_loss(DevScalars* sc, int R, double alpha, double beta, double gamma, int use_tv, double* loss_out) {
    double s_corr = 0.0, s_con = 0.0;
    const double zc = sc->zero[0].contrast;
    for (int q = 0; q < R; ++q) {
        const Stats st = sc->ref[q];
        const double w = sc->weights[q];
        s_corr += (w * (-st.mse)) / ((-sc->zero[q].mse) + kEps);                         // losses.py:176
        s_con += (w * st.contrast) / (zc + kEps);                                        // losses.py:177
    }
    const double mean_rel_corr = s_corr / R, mean_rel_contrast = s_con / R;
    const double tv = use_tv ? sc->tv_sum / (sc->tv_cnt + kEps) : 0.0;                  // regularizers.py:31-36, losses.py:171
    const double loss = (alpha * (mean_rel_contrast * (-1.0)) + beta * (mean_rel_corr * (-1.0))) + (gamma * tv + 0.0);
    sc->loss = loss; sc->mean_rel_corr = mean_rel_corr; sc->mean_rel_contrast = mean_rel_contrast;
    sc->mean_rel_div = 0.0; sc->tv = tv;
    if (loss_out != nullptr) *loss_out = loss;
}

__device__ __forceinline__ double shfl_up_d(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_dn_d(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

__device__ __forceinline__ void image_stats_body(const ImageStatsArgs& A) {
    __shared__ StatsTail S;
    const int H = A.H, W = A.W, R = A.R;
    const int HW = H * W;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int strips = image_stats_strips(W), bands = image_stats_bands(H);
    const int per_img = strips * bands, cpi = image_stats_ctas(H, W);
    const int G = gridDim.x, b = blockIdx.x;
    // programmatic dependent launch: the next kernel of the evaluation may be scheduled as soon as every CTA of this one is resident;
    // this kernel itself may have been scheduled while the splat was still running - nothing is read or written before the wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (A.skip != nullptr && *reinterpret_cast<const volatile int*>(A.skip) != 0) return;
    // accumulators of the event backward pass
    if (A.zero_buf != nullptr)
        for (int k = b * kS2NT + tid; k < A.n_zero; k += G * kS2NT) A.zero_buf[k] = 0.0;
    if (A.zero_buf2 != nullptr)
        for (int k = b * kS2NT + tid; k < A.n_zero2; k += G * kS2NT) A.zero_buf2[k] = 0.0;

    for (int cta = b; cta < R * cpi; cta += G) {
        const int r = cta / cpi, ci = cta - r * cpi;
        const int rem = ci * kS2Warps + wid;                      // work item of this warp inside image r (idle beyond per_img)
        FusedAcc acc;
        acc.init();
        if (rem < per_img) {
        const int band = rem / strips, strip = rem - band * strips;
        const int x = strip * kS2Cols - 2 + lane;                 // this lane's column (lanes 0, 1, 30, 31: halo)
        const int y0 = band * kS2Rows, y1 = min(H, y0 + kS2Rows);
        const bool xin = x >= 0 && x < W;
        const bool own_col = lane >= 2 && lane < 2 + kS2Cols && xin;
        const unsigned long long* Fr = A.fix + r * HW + (xin ? x : 0);
        const double* Er = A.edges + r * HW + (xin ? x : 0);
        double* Ir = A.iwe + r * HW;
        float* Ar = A.adj32 + r * HW;
        int cnt_mn = 0, cnt_mx = 0;                    // tie counts of the running min / max (integers: branch-free update)
        // three-row windows: image value and horizontal difference of input rows yi-2, yi-1 (yi: the row being consumed);
        // Scharr x differences and Scharr y of rows yc-2, yc-1 (yc = yi - 1: the row whose Scharr pair is produced)
        double I2 = 0.0, I1 = 0.0, Dx2 = 0.0, Dx1 = 0.0, E1 = 0.0;
        float Ex2 = 0.f, Ex1 = 0.f, gy2 = 0.f, gy1 = 0.f;
        // load groups: rows base .. base + kS2Pre - 1, one group ahead
        unsigned long long fq[kS2Pre];
        double eq[kS2Pre];
        auto load_group = [&](int base) {
#pragma unroll
            for (int u = 0; u < kS2Pre; ++u) {
                const int y = base + u;
                const bool ok = xin && y >= 0 && y < H;
                fq[u] = ok ? __ldcg(Fr + y * W) : 0ull;
                eq[u] = (ok && y >= y0 && y < y1) ? __ldg(Er + y * W) : 0.0;
            }
        };
        load_group(y0 - 2);
        for (int base = y0 - 2; base <= y1 + 1; base += kS2Pre) {
            unsigned long long fc[kS2Pre];
            double ec[kS2Pre];
#pragma unroll
            for (int u = 0; u < kS2Pre; ++u) { fc[u] = fq[u]; ec[u] = eq[u]; }
            if (base + kS2Pre <= y1 + 1) load_group(base + kS2Pre);
#pragma unroll
            for (int u = 0; u < kS2Pre; ++u) {
                const int yi = base + u;
                if (yi > y1 + 1) break;                                     // warp-uniform
                // fixed point -> float64: sums are far below 2^52, (2^52 | v) reinterpreted as float64 is exactly 2^52 + v
                const double I0 = __dsub_rn(__longlong_as_double(0x4330000000000000ll | (long long)fc[u]), 4503599627370496.0) * kFixToIwe;
                const double Dx0 = __dsub_rn(shfl_dn_d(I0), shfl_up_d(I0));   // I[yi][x+1] - I[yi][x-1] (lanes 0 / 31: unused garbage)
                // ---- Scharr pair of row yc = yi - 1 (zero outside the image: the 'same' output only exists inside)
                const int yc = yi - 1;
                const double Dy = __dsub_rn(I0, I2);                          // I[yc+1][x] - I[yc-1][x]
                const double Dyl = shfl_up_d(Dy), Dyr = shfl_dn_d(Dy);
                double gx = 0.0, gy = 0.0;
                if (xin && yc >= 0 && yc < H) {
                    gx = __dadd_rn(__dadd_rn(__dmul_rn(3.0, Dx0), __dmul_rn(10.0, Dx1)), __dmul_rn(3.0, Dx2));
                    gy = __dadd_rn(__dadd_rn(__dmul_rn(3.0, Dyr), __dmul_rn(10.0, Dy)), __dmul_rn(3.0, Dyl));
                }
                if (own_col && yc >= y0 && yc < y1) {                         // statistics and image value of an own pixel
                    const double I = I1;
                    Ir[yc * W + x] = I;
                    acc.sq += gx * gx + gy * gy; acc.sI += I; acc.sI2 += I * I; acc.sEI += E1 * I;
                    cnt_mn = (I < acc.mn) ? 1 : cnt_mn + (I == acc.mn ? 1 : 0);
                    cnt_mx = (I > acc.mx) ? 1 : cnt_mx + (I == acc.mx ? 1 : 0);
                    acc.mn = fmin(acc.mn, I);
                    acc.mx = fmax(acc.mx, I);
                }
                // ---- adjoint of the Scharr pair at row ya = yc - 1 from the float32 pair of rows yc-2, yc-1, yc
                const float gx0 = (float)gx, gy0 = (float)gy;
                const float Ex0 = __fsub_rn(__shfl_up_sync(0xffffffffu, gx0, 1), __shfl_down_sync(0xffffffffu, gx0, 1));   // gx[x-1] - gx[x+1]
                const float Fy = __fsub_rn(gy2, gy0);                          // gy[ya-1] - gy[ya+1]
                const float Fyl = __shfl_up_sync(0xffffffffu, Fy, 1), Fyr = __shfl_down_sync(0xffffffffu, Fy, 1);
                const int ya = yc - 1;
                if (own_col && ya >= y0 && ya < y1) {
                    const float ax = __fadd_rn(__fadd_rn(__fmul_rn(3.f, Ex2), __fmul_rn(10.f, Ex1)), __fmul_rn(3.f, Ex0));
                    const float ay = __fadd_rn(__fadd_rn(__fmul_rn(3.f, Fyl), __fmul_rn(10.f, Fy)), __fmul_rn(3.f, Fyr));
                    const float adj = __fadd_rn(ax, ay);
                    Ar[ya * W + x] = adj;
                }
                // shift the windows
                I2 = I1; I1 = I0; Dx2 = Dx1; Dx1 = Dx0; E1 = ec[u];
                Ex2 = Ex1; Ex1 = Ex0; gy2 = gy1; gy1 = gy0;
            }
        }
        acc.cmn = (double)cnt_mn; acc.cmx = (double)cnt_mx;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) acc.merge(acc.shfl_xor(o));
        // one record per CTA: the warps' partials merged in warp order
        __syncthreads();                               // S.wacc of the previous round has been read
        if (lane == 0) {
            double* d = S.wacc[wid];
            d[0] = acc.sq; d[1] = acc.sI; d[2] = acc.sI2; d[3] = acc.sEI; d[4] = acc.mn; d[5] = acc.cmn; d[6] = acc.mx; d[7] = acc.cmx;
        }
        __syncthreads();
        if (tid == 0) {
            FusedAcc a;
            a.init();
            for (int k = 0; k < kS2Warps; ++k) {
                FusedAcc o;
                o.sq = S.wacc[k][0]; o.sI = S.wacc[k][1]; o.sI2 = S.wacc[k][2]; o.sEI = S.wacc[k][3];
                o.mn = S.wacc[k][4]; o.cmn = S.wacc[k][5]; o.mx = S.wacc[k][6]; o.cmx = S.wacc[k][7];
                a.merge(o);
            }
            double* d = A.part + (int64_t)cta * kFPart;
            d[0] = a.sq; d[1] = a.sI; d[2] = a.sI2; d[3] = a.sEI; d[4] = a.mn; d[5] = a.cmn; d[6] = a.mx; d[7] = a.cmx;
        }
    }
    // ---- tail: the last CTA to finish reduces the records (one warp per reference image) --------------------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) S.is_last = atomicAdd(&A.sc->counters[kStatsTicket], 1u) == (unsigned)(G - 1);
    __syncthreads();
    if (!S.is_last) return;
    __threadfence();
    stats_tail<kS2NT>(A.part, A.sc, R, cpi, HW, A.alpha, A.beta);
    if (tid == 0) A.sc->counters[kStatsTicket] = 0u;
}

__global__ void __launch_bounds__(kS2NT, EINCM_S2_MINB)
k_image_stats(const ImageStatsArgs A) { image_stats_body(A); }

// batched form: blockIdx.y = window, one argument record per window in device memory
__global__ void __launch_bounds__(kS2NT, EINCM_S2_MINB)
k_image_stats_b(const ImageStatsArgs* __restrict__ args, const int* __restrict__ order) {
    __shared__ ImageStatsArgs sA;
    const int bw = batch_window(order);
    if (bw < 0) return;
    load_args(sA, args + bw);
    image_stats_body(sA);
}

// ---- k_image_grad ----------------------------------------------------------------------------------------------------------
// d loss / d IWE_r = cA_r * adjoint(Gx, Gy) + gN / D + (tie-split cotangents of min and max)   with   gN = cB_r (E - (I - m) / D)
// (reverse mode of contrast_objectives.py:22-25, img_utils.py:24-25, correlation_objectives.py:25-26; min / max cotangents
// split evenly among ties like jnp.min / jnp.max).  Pointwise: four cells per thread and round, loads batched.  Also clears the
// fixed-point images (every reader has finished: kernel boundary).
struct ImageGradArgs {
    unsigned long long* fix;        // [R][H*W] cleared (null: left alone)
    const double* edges;
    const double* iwe;
    const float* adj32;
    DevScalars* sc;
    const int* skip;                // non-zero: return at once (unrolled solve graphs), or null
    int publish;                    // != 0: the grid has one CTA more than the pointwise pass needs; it publishes the loss (publish_loss)
    int use_tv;
    double* loss_out;
    double alpha, beta, gamma;
    double* dldi;                   // [R][H*W] float64 d loss / d IWE (debug tap) or null
    float* dldi32;                  // [R][H*W] float32 d loss / d IWE / (2 pi) or null
    int HW, R;
    int want_grad;                  // 0: only clear the fixed-point images
};

__device__ __forceinline__ void image_grad_body(const ImageGradArgs& A) {
    // programmatic dependent launch (no-ops when launched plainly): let the next kernel be scheduled, wait for the image statistics
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (A.skip != nullptr && *reinterpret_cast<const volatile int*>(A.skip) != 0) return;
    __shared__ double s_mn[kCoopMaxRefs], s_mx[kCoopMaxRefs], s_iD[kCoopMaxRefs], s_cA[kCoopMaxRefs], s_cB[kCoopMaxRefs], s_tm[kCoopMaxRefs], s_tM[kCoopMaxRefs];
    // The last CTA of a publishing launch only evaluates the loss from the statistics (a serial chain of float64 divisions that nothing
    // before k_theta_grad waits for) and leaves; the pointwise pass belongs to the other CTAs.
    if (A.publish && blockIdx.x == gridDim.x - 1) {
        if (threadIdx.x == 0) publish_loss(A.sc, A.R, A.alpha, A.beta, A.gamma, A.use_tv, A.loss_out);
        return;
    }
    if (threadIdx.x < A.R && A.want_grad) {
        const Stats st = A.sc->ref[threadIdx.x];
        s_mn[threadIdx.x] = st.mn; s_mx[threadIdx.x] = st.mx; s_iD[threadIdx.x] = 1.0 / st.D;
        s_cA[threadIdx.x] = A.sc->coefA[threadIdx.x]; s_cB[threadIdx.x] = A.sc->coefB[threadIdx.x];
        const double g_M = -st.s2 / (st.D * st.D);
        const double g_m = -st.s1 / st.D + st.s2 / (st.D * st.D);
        s_tm[threadIdx.x] = g_m / st.cnt_min; s_tM[threadIdx.x] = g_M / st.cnt_max;
    }
    __syncthreads();
    const int T = (A.publish ? gridDim.x - 1 : gridDim.x) * blockDim.x;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < A.R; ++r) {
        // 1 / D is a per-image constant: the two divisions of the reference formula become multiplications (one rounding more)
        const double mn = s_mn[r], mx = s_mx[r], iD = s_iD[r], cA = s_cA[r], cBD = s_cB[r] * s_iD[r], tm = s_tm[r], tM = s_tM[r];
        const double* Er = A.edges + (size_t)r * A.HW;
        const double* Ir = A.iwe + (size_t)r * A.HW;
        const float* Ar = A.adj32 + (size_t)r * A.HW;
        for (int base = gt; base < A.HW; base += 4 * T) {
            double I[4], E[4];
            float a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int p = min(base + u * T, A.HW - 1);
                if (A.want_grad) { I[u] = __ldcg(Ir + p); E[u] = __ldg(Er + p); a[u] = __ldcg(Ar + p); }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int p = base + u * T;
                if (p < A.HW) {
                    if (A.fix != nullptr) A.fix[(size_t)r * A.HW + p] = 0ull;
                    if (A.want_grad) {
                        double out = cA * (double)a[u] + cBD * (E[u] - (I[u] - mn) * iD);
                        if (I[u] == mn) out += tm;
                        if (I[u] == mx) out += tM;
                        if (A.dldi != nullptr) A.dldi[(size_t)r * A.HW + p] = out;
                        if (A.dldi32 != nullptr) A.dldi32[(size_t)r * A.HW + p] = (float)(out * kInv2Pi);
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_image_grad(const ImageGradArgs A) { image_grad_body(A); }

__global__ void __launch_bounds__(256)
k_image_grad_b(const ImageGradArgs* __restrict__ args, const int* __restrict__ order) {
    __shared__ ImageGradArgs sA;
    const int bw = batch_window(order);
    if (bw < 0) return;
    load_args(sA, args + bw);
    image_grad_body(sA);
}

// per-window: sum E_r and sum E_r^2 (deterministic single-CTA-per-reference reduction; once per window)
__global__ void __launch_bounds__(1024)
k_edge_sums(const double* __restrict__ edges, int64_t HW, DevScalars* sc) {
    __shared__ double sh[32];
    const int r = blockIdx.x;
    double s = 0.0, s2 = 0.0;
    for (int64_t p = threadIdx.x; p < HW; p += blockDim.x) {
        const double e = edges[r * HW + p];
        s += e; s2 += e * e;
    }
    s = block_reduce<1024>(s, OpSum(), sh);
    s2 = block_reduce<1024>(s2, OpSum(), sh);
    if (threadIdx.x == 0) { sc->sumE[r] = s; sc->sumE2[r] = s2; }
}

}  // namespace eincm
