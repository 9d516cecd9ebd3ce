// Shared device helpers and the device-side scalar block of an EINCM plan.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define EINCM_MAX_REFS 8

namespace eincm {

constexpr double kEps = 2.220446049250313e-16;     // sys.float_info.epsilon (reference src/eincm/losses.py:24)
constexpr double kLog2Pi = 1.8378770664093453;     // math.log(2*pi)
constexpr double kInv2Pi = 0.15915494309189535;    // 1/(2*pi)

// Per-image statistics produced by the image-space kernels (one slot per reference time, plus the zero-IWE).
struct Stats {
    double contrast;   // mean(Gx^2 + Gy^2)              contrast_objectives.py:22-25
    double mn, mx, D;  // min, max, max - min + eps      img_utils.py:25
    double mse;        // mean((E - N)^2)                correlation_objectives.py:25-26
    double s1, s2;     // sum gN, sum gN*(I - mn)        (min-max normalise backward)
    double cnt_min, cnt_max;   // tie counts for the min / max cotangent split
    double div;        // iwe_divergence                 event_collapse_objectives.py:8-20
};

struct DevScalars {
    Stats ref[EINCM_MAX_REFS];    // warped IWE_r
    Stats zero[EINCM_MAX_REFS];   // zero-IWE against edge_r (contrast/min/max/div live in zero[0])
    double weights[EINCM_MAX_REFS];   // multi-reference weights (losses.py:39-46)
    double tv_sum, tv_cnt;        // regularizers.py:31-36 numerator / non-zero-gradient pixel count
    double loss, mean_rel_corr, mean_rel_contrast, mean_rel_div, tv, dalpha;
    double coefA[EINCM_MAX_REFS]; // a_r * 2/HW   (contrast cotangent scale)
    double coefB[EINCM_MAX_REFS]; // b_r * -2/HW  (correlation cotangent scale)
    double coefD[EINCM_MAX_REFS]; // d_r / HW     (divergence cotangent scale)
    double sumE[EINCM_MAX_REFS], sumE2[EINCM_MAX_REFS];   // per-window sums of edge_r and edge_r^2 (fused image pass)
    double eval_seq;              // evaluations delivered to the host so far (incremented by the kernel that delivers a result)
    unsigned int counters[16];    // "last block done" tickets
    int error_flag;               // set by kernels on invalid input (event outside the sensor)
};

struct RefTimes { double t[EINCM_MAX_REFS]; };

__device__ __forceinline__ int linear_tid() { return threadIdx.y * blockDim.x + threadIdx.x; }

// Sum / min / max over a thread block of NT threads (NT multiple of 32, <= 1024).  Result valid in thread 0.
template <int NT, typename Op>
__device__ __forceinline__ double block_reduce(double v, Op op, double* sh /* NT/32 doubles */) {
    const int tid = linear_tid();
#pragma unroll
    for (int o = 16; o; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((tid & 31) == 0) sh[tid >> 5] = v;
    __syncthreads();
    double acc = sh[0];
    if (tid == 0) {
        for (int k = 1; k < NT / 32; ++k) acc = op(acc, sh[k]);   // fixed order: deterministic
    }
    __syncthreads();   // sh may be reused by the caller
    return acc;
}

struct OpSum { __device__ __forceinline__ double operator()(double a, double b) const { return a + b; } };
struct OpMin { __device__ __forceinline__ double operator()(double a, double b) const { return fmin(a, b); } };
struct OpMax { __device__ __forceinline__ double operator()(double a, double b) const { return fmax(a, b); } };

// Scharr gradients in the canonical summation order of oracle.sobel_scharr_optimized_image_grads
// (difference first, multiplies and adds NOT contracted into FMAs so exact zeros / signs match the oracle):
//   Gx = (3*(I[i+1,j+1]-I[i+1,j-1]) + 10*(I[i,j+1]-I[i,j-1])) + 3*(I[i-1,j+1]-I[i-1,j-1])
//   Gy = (3*(I[i+1,j+1]-I[i-1,j+1]) + 10*(I[i+1,j]-I[i-1,j])) + 3*(I[i+1,j-1]-I[i-1,j-1])
// p points at I[i][j] inside a shared-memory tile with row pitch `pitch` (reference src/utils/img_utils.py:414-425).
// a..f, u, v = I[i+1,j+1], I[i+1,j-1], I[i,j+1], I[i,j-1], I[i-1,j+1], I[i-1,j-1], I[i+1,j], I[i-1,j]
__device__ __forceinline__ void scharr_vals(double a, double b, double c, double d, double e, double f, double u, double v,
                                            double& gx, double& gy) {
    gx = __dadd_rn(__dadd_rn(__dmul_rn(3.0, __dsub_rn(a, b)), __dmul_rn(10.0, __dsub_rn(c, d))), __dmul_rn(3.0, __dsub_rn(e, f)));
    gy = __dadd_rn(__dadd_rn(__dmul_rn(3.0, __dsub_rn(a, e)), __dmul_rn(10.0, __dsub_rn(u, v))), __dmul_rn(3.0, __dsub_rn(b, f)));
}

__device__ __forceinline__ void scharr_at(const double* p, int pitch, double& gx, double& gy) {
    scharr_vals(p[pitch + 1], p[pitch - 1], p[1], p[-1], p[-pitch + 1], p[-pitch - 1], p[pitch], p[-pitch], gx, gy);
}

// Adjoint of the Scharr pair ('same' correlation with the same kernels), canonical order of oracle._scharr_adjoint.
// xu/xm/xd point at the cotangent of Gx at column j of rows i-1 / i / i+1 (yu/ym/yd likewise for Gy).
template <typename T>
__device__ __forceinline__ double scharr_adjoint_rows(const T* xu, const T* xm, const T* xd, const T* yu, const T* yd) {
    const double ax = __dadd_rn(__dadd_rn(__dmul_rn(3.0, __dsub_rn((double)xu[-1], (double)xu[1])),
                                          __dmul_rn(10.0, __dsub_rn((double)xm[-1], (double)xm[1]))),
                                __dmul_rn(3.0, __dsub_rn((double)xd[-1], (double)xd[1])));
    const double ay = __dadd_rn(__dadd_rn(__dmul_rn(3.0, __dsub_rn((double)yu[-1], (double)yd[-1])),
                                          __dmul_rn(10.0, __dsub_rn((double)yu[0], (double)yd[0]))),
                                __dmul_rn(3.0, __dsub_rn((double)yu[1], (double)yd[1])));
    return __dadd_rn(ax, ay);
}

// px / py point at the cotangents of Gx / Gy at [i][j] in shared-memory tiles with row pitch `pitch`.
__device__ __forceinline__ double scharr_adjoint_at(const double* px, const double* py, int pitch) {
    return scharr_adjoint_rows(px - pitch, px, px + pitch, py - pitch, py + pitch);
}

// convolve(a, DIV_KERN, 'same') in the canonical order of oracle.div_kern_conv (event_collapse_objectives.py:14-16)
__device__ __forceinline__ double divk_at(const double* p, int pitch) {
    const double corners = __dadd_rn(__dadd_rn(__dadd_rn(p[pitch + 1], p[pitch - 1]), p[-pitch + 1]), p[-pitch - 1]);
    const double edges = __dadd_rn(__dadd_rn(__dadd_rn(p[pitch], p[1]), p[-1]), p[-pitch]);
    return __dadd_rn(__dmul_rn(corners, 1.0 / 12.0), __dmul_rn(edges, 1.0 / 6.0));
}

__device__ __forceinline__ double sign_of(double v) { return (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : 0.0); }

// Index rule of `frame.at[r, c].add(v, mode='drop')` (reference src/utils/event_utils.py:59):
// negative indices in [-N, -1] wrap (NumPy-style normalisation), anything still outside [0, N) is dropped.
// float32 form for cotangents that only exist in float32 (k_image_stats): same order of operations
__device__ __forceinline__ float scharr_adjoint_rows_f32(const float* xu, const float* xm, const float* xd, const float* yu, const float* yd) {
    const float ax = __fadd_rn(__fadd_rn(__fmul_rn(3.f, __fsub_rn(xu[-1], xu[1])), __fmul_rn(10.f, __fsub_rn(xm[-1], xm[1]))),
                               __fmul_rn(3.f, __fsub_rn(xd[-1], xd[1])));
    const float ay = __fadd_rn(__fadd_rn(__fmul_rn(3.f, __fsub_rn(yu[-1], yd[-1])), __fmul_rn(10.f, __fsub_rn(yu[0], yd[0]))),
                               __fmul_rn(3.f, __fsub_rn(yu[1], yd[1])));
    return __fadd_rn(ax, ay);
}

template <bool WRAP>
__device__ __forceinline__ bool drop_index(int& r, int& c, int H, int W) {
    if (WRAP) {
        if (r < 0) r += H;
        if (c < 0) c += W;
    }
    return (r >= 0) & (r < H) & (c >= 0) & (c < W);
}

}  // namespace eincm
