// Evaluation metrics computed from the images / flow fields an evaluation leaves on the device:
//   flow warp loss   var(IWE_r) / var(zero-IWE)        reference src/eincm/objectives/contrast_metrics.py:6-17, losses.py:84
//   theta divergence mean |K*a + K*b + K*c + K*d|       reference src/eincm/regularizers.py:41-58
//   sparse flow error (AEE, AREE, N-pixel error percentages, counts)   reference src/evaluations/flow_eval.py:14-75
// Not on the optimisation hot path (one call per solved window): simple deterministic reductions.
#pragma once
#include "common.cuh"

namespace eincm {

// np.var of n_img images laid out [n_img][n]: two passes inside one CTA per image (mean, then mean squared deviation)
__global__ void __launch_bounds__(1024)
k_image_var(const double* __restrict__ imgs, int64_t n, double* __restrict__ var_out, double* __restrict__ mean_out) {
    __shared__ double sh[32];
    __shared__ double s_mean;
    const double* img = imgs + (int64_t)blockIdx.x * n;
    double s = 0.0;
    for (int64_t p = threadIdx.x; p < n; p += blockDim.x) s += img[p];
    s = block_reduce<1024>(s, OpSum(), sh);
    if (threadIdx.x == 0) s_mean = s / (double)n;
    __syncthreads();
    const double m = s_mean;
    double q = 0.0;
    for (int64_t p = threadIdx.x; p < n; p += blockDim.x) { const double d = img[p] - m; q += d * d; }
    q = block_reduce<1024>(q, OpSum(), sh);
    if (threadIdx.x == 0) { var_out[blockIdx.x] = q / (double)n; if (mean_out) mean_out[blockIdx.x] = m; }
}

// sums n_cols columns of part[n_rows][n_cols] in row order (fixed order: deterministic); one thread per column
__global__ void k_sum_rows(const double* __restrict__ part, int n_rows, int n_cols, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    double s = 0.0;
    for (int r = 0; r < n_rows; ++r) s += part[(size_t)r * n_cols + c];
    out[c] = s;
}

// ---- theta divergence (regularizers.py:41-58): a,b = Scharr(theta_x), c,d = Scharr(theta_y), s = ((K*a + K*b) + K*c) + K*d with the
// 3x3 divergence kernel K (symmetric: convolution == correlation), all 'same' with zero padding; result = mean |s|.
// One CTA per 32x8 tile: flow with halo 2 -> gradients with halo 1 -> s.  part[block] = sum |s| over the tile.
constexpr int kEvTX = 32, kEvTY = 8;

__device__ __forceinline__ double div_kern_at(const double* p, int pitch) {
    const double corners = ((p[pitch + 1] + p[pitch - 1]) + p[-pitch + 1]) + p[-pitch - 1];      // oracle.div_kern_conv order
    const double edges = ((p[pitch] + p[1]) + p[-1]) + p[-pitch];
    return corners * (1.0 / 12.0) + edges * (1.0 / 6.0);
}

__global__ void __launch_bounds__(kEvTX * kEvTY)
k_theta_divergence(const double2* __restrict__ theta_full, int H, int W, double* __restrict__ part) {
    constexpr int PW = kEvTX + 4, PH = kEvTY + 4, GW = kEvTX + 2, GH = kEvTY + 2;
    __shared__ double fx[PH][PW], fy[PH][PW];
    __shared__ double ga[GH][GW], gb[GH][GW], gc[GH][GW], gd[GH][GW];
    __shared__ double red[8];
    const int tid = linear_tid();
    const int x0 = blockIdx.x * kEvTX, y0 = blockIdx.y * kEvTY;
    for (int k = tid; k < PW * PH; k += kEvTX * kEvTY) {
        const int ly = k / PW, lx = k % PW;
        const int y = y0 + ly - 2, x = x0 + lx - 2;
        double2 t = make_double2(0.0, 0.0);
        if (x >= 0 && x < W && y >= 0 && y < H) t = theta_full[(size_t)y * W + x];
        fx[ly][lx] = t.x; fy[ly][lx] = t.y;
    }
    __syncthreads();
    for (int k = tid; k < GW * GH; k += kEvTX * kEvTY) {
        const int ly = k / GW, lx = k % GW;
        const int y = y0 + ly - 1, x = x0 + lx - 1;
        double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
        if (x >= 0 && x < W && y >= 0 && y < H) {            // 'same' output only exists inside the image
            scharr_at(&fx[ly + 1][lx + 1], PW, a, b);
            scharr_at(&fy[ly + 1][lx + 1], PW, c, d);
        }
        ga[ly][lx] = a; gb[ly][lx] = b; gc[ly][lx] = c; gd[ly][lx] = d;
    }
    __syncthreads();
    double v = 0.0;
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x < W && y < H) {
        const int ly = threadIdx.y + 1, lx = threadIdx.x + 1;
        const double s = ((div_kern_at(&ga[ly][lx], GW) + div_kern_at(&gb[ly][lx], GW)) + div_kern_at(&gc[ly][lx], GW)) + div_kern_at(&gd[ly][lx], GW);
        v = fabs(s);
    }
    v = block_reduce<kEvTX * kEvTY>(v, OpSum(), red);
    if (tid == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = v;
}

// ---- sparse flow error (flow_eval.py:14-75) --------------------------------------------------------------------------------
// pred_flow [H*W][2] (multiplied by pred_mult[p] != 0 when pred_mult is given: per_pix_theta_to_flow, theta_utils.py:40-73),
// gt_flow [H*W][2], event_mask [H*W] or null.  part[block][kFlowErrCols]:
//   0 n_pred, 1 n_gt, 2 n_ee, 3 sum ee, 4 sum ree, 5..10 count(ee > 1, 2, 3, 5, 10, 20)
constexpr int kFlowErrCols = 11;

__global__ void __launch_bounds__(256)
k_flow_error(const double2* __restrict__ pred_flow, const uint8_t* __restrict__ pred_mult, const double2* __restrict__ gt_flow,
             const uint8_t* __restrict__ event_mask, int64_t n, double* __restrict__ part) {
    __shared__ double red[8];
    double acc[kFlowErrCols];
#pragma unroll
    for (int k = 0; k < kFlowErrCols; ++k) acc[k] = 0.0;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        double2 pr = pred_flow[p];
        if (pred_mult != nullptr) { const double m = pred_mult[p] ? 1.0 : 0.0; pr.x *= m; pr.y *= m; }
        const double2 gt = gt_flow[p];
        // ~isinf & ~isinf & (norm > 0): a NaN component fails the norm test                                     flow_eval.py:32-36
        bool mp = !isinf(pr.x) && !isinf(pr.y) && (sqrt(pr.x * pr.x + pr.y * pr.y) > 0.0);
        if (event_mask != nullptr) mp = mp && event_mask[p] != 0;                                             // :38-39
        const double ng = sqrt(gt.x * gt.x + gt.y * gt.y);
        const bool mg = !isinf(gt.x) && !isinf(gt.y) && (ng > 0.0);                                           // :42-46
        acc[0] += mp ? 1.0 : 0.0;
        acc[1] += mg ? 1.0 : 0.0;
        if (mp && mg) {                                                                                       // :49
            const double dx = pr.x - gt.x, dy = pr.y - gt.y;
            const double ee = sqrt(dx * dx + dy * dy);                                                        // :56
            acc[2] += 1.0;
            acc[3] += ee;
            acc[4] += ee / (ng + kEps);                                                                       // :59
            acc[5] += ee > 1.0 ? 1.0 : 0.0; acc[6] += ee > 2.0 ? 1.0 : 0.0; acc[7] += ee > 3.0 ? 1.0 : 0.0;       // :72-74
            acc[8] += ee > 5.0 ? 1.0 : 0.0; acc[9] += ee > 10.0 ? 1.0 : 0.0; acc[10] += ee > 20.0 ? 1.0 : 0.0;
        }
    }
#pragma unroll
    for (int k = 0; k < kFlowErrCols; ++k) {
        const double v = block_reduce<256>(acc[k], OpSum(), red);
        if (threadIdx.x == 0) part[(size_t)blockIdx.x * kFlowErrCols + k] = v;
    }
}

}  // namespace eincm
