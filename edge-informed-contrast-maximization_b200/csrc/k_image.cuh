// Image-space kernels on the R images of warped events (all float64, images are L2-resident):
//   A: Scharr contrast, min, max                      (contrast_objectives.py:13-26, img_utils.py:24-25,414-425)
//   B: MSE against the edge image of the min-max-normalised IWE, and the sums its backward needs
//                                                      (correlation_objectives.py:12-27, losses.py:62-67)
//   C: d loss / d IWE  (transpose Scharr of the contrast term + normalise/MSE chain incl. min/max tie split)
//   D1/D2: optional IWE-divergence objective (delta != 0; event_collapse_objectives.py:8-20) forward / backward
// Global reductions are two-level and deterministic: per-CTA partials, combined in a fixed order by the
// last CTA to finish.
#pragma once
#include "common.cuh"

namespace eincm {

constexpr int kImgTX = 32, kImgTY = 8, kImgNT = kImgTX * kImgTY;

// Finishes a set of per-block partial sums: returns true in the last block to arrive.
__device__ __forceinline__ bool last_block_ticket(unsigned int* counter, unsigned int total_blocks, bool* sh_flag) {
    if (linear_tid() == 0) {
        __threadfence();
        *sh_flag = (atomicAdd(counter, 1u) == total_blocks - 1u);
    }
    __syncthreads();
    return *sh_flag;
}

// load a (kImgTY + 2*HALO) x (kImgTX + 2*HALO) tile of img around the CTA's tile, zero outside the image
template <int HALO>
__device__ __forceinline__ void load_tile(const double* __restrict__ img, int H, int W, int x0, int y0,
                                          double (*tile)[kImgTX + 2 * HALO]) {
    constexpr int PW = kImgTX + 2 * HALO, PH = kImgTY + 2 * HALO;
    for (int k = linear_tid(); k < PW * PH; k += kImgNT) {
        const int ly = k / PW, lx = k % PW;
        const int y = y0 + ly - HALO, x = x0 + lx - HALO;
        tile[ly][lx] = (x >= 0 && x < W && y >= 0 && y < H) ? img[y * W + x] : 0.0;
    }
}

// ---- A ---------------------------------------------------------------------------------------------------
// grid (tiles_x, tiles_y, n_img); part: [3][n_img][nb]; writes stats[z].contrast/mn/mx/D
__global__ void __launch_bounds__(kImgNT)
k_img_A(const double* __restrict__ imgs, int H, int W, int nb, double* __restrict__ part, Stats* __restrict__ stats,
        unsigned int* __restrict__ counter) {
    __shared__ double tile[kImgTY + 2][kImgTX + 2];
    __shared__ double red[8];
    __shared__ bool last;
    const int z = blockIdx.z, n_img = gridDim.z;
    const double* img = imgs + (int64_t)z * H * W;
    const int x0 = blockIdx.x * kImgTX, y0 = blockIdx.y * kImgTY;
    load_tile<1>(img, H, W, x0, y0, tile);
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    double sq = 0.0, mn = INFINITY, mx = -INFINITY;
    if (x < W && y < H) {
        double gx, gy;
        scharr_at(&tile[threadIdx.y + 1][threadIdx.x + 1], kImgTX + 2, gx, gy);
        sq = gx * gx + gy * gy;
        mn = mx = tile[threadIdx.y + 1][threadIdx.x + 1];
    }
    sq = block_reduce<kImgNT>(sq, OpSum(), red);
    mn = block_reduce<kImgNT>(mn, OpMin(), red);
    mx = block_reduce<kImgNT>(mx, OpMax(), red);
    const int b = blockIdx.y * gridDim.x + blockIdx.x;
    if (linear_tid() == 0) {
        part[(0 * n_img + z) * nb + b] = sq;
        part[(1 * n_img + z) * nb + b] = mn;
        part[(2 * n_img + z) * nb + b] = mx;
    }
    if (!last_block_ticket(counter, (unsigned)(nb * n_img), &last)) return;
    for (int r = 0; r < n_img; ++r) {
        double s = 0.0, lo = INFINITY, hi = -INFINITY;
        for (int k = linear_tid(); k < nb; k += kImgNT) {
            s += __ldcg(&part[(0 * n_img + r) * nb + k]);
            lo = fmin(lo, __ldcg(&part[(1 * n_img + r) * nb + k]));
            hi = fmax(hi, __ldcg(&part[(2 * n_img + r) * nb + k]));
        }
        s = block_reduce<kImgNT>(s, OpSum(), red);
        lo = block_reduce<kImgNT>(lo, OpMin(), red);
        hi = block_reduce<kImgNT>(hi, OpMax(), red);
        if (linear_tid() == 0) {
            stats[r].contrast = s / ((double)H * (double)W);
            stats[r].mn = lo; stats[r].mx = hi;
            stats[r].D = (hi - lo) + kEps;                       // img_utils.py:25
        }
    }
    if (linear_tid() == 0) *counter = 0;
}

// ---- B ---------------------------------------------------------------------------------------------------
// grid (tiles, 1, R).  img_stride == 0: every reference uses the same image / stats slot 0 (zero-IWE pass).
// gN = coefB[r] * (E - N) (+ gNdiv);  part: [5][R][nb] = mse, s1, s2, cnt_min, cnt_max
__global__ void __launch_bounds__(kImgNT)
k_img_B(const double* __restrict__ imgs, int64_t img_stride, const double* __restrict__ edges, const double* __restrict__ gNdiv,
        int64_t HW, int nb, const Stats* stats_in, int stats_in_stride, const double* __restrict__ coefB,
        double* __restrict__ part, Stats* stats_out, unsigned int* __restrict__ counter) {
    __shared__ double red[8];
    __shared__ bool last;
    const int r = blockIdx.z, R = gridDim.z;
    const double* img = imgs + r * img_stride;
    const double* E = edges + r * HW;
    const Stats st = stats_in[r * stats_in_stride];
    const double cb = coefB != nullptr ? coefB[r] : 0.0;
    double mse = 0.0, s1 = 0.0, s2 = 0.0, cmin = 0.0, cmax = 0.0;
    for (int64_t p = blockIdx.x * (int64_t)kImgNT + linear_tid(); p < HW; p += (int64_t)gridDim.x * kImgNT) {
        const double I = img[p];
        const double c = I - st.mn;
        const double N = c / st.D;
        const double d = E[p] - N;
        mse += d * d;
        double gN = cb * d;
        if (gNdiv != nullptr) gN += gNdiv[r * HW + p];
        s1 += gN;
        s2 += gN * c;
        cmin += (I == st.mn) ? 1.0 : 0.0;
        cmax += (I == st.mx) ? 1.0 : 0.0;
    }
    mse = block_reduce<kImgNT>(mse, OpSum(), red);
    s1 = block_reduce<kImgNT>(s1, OpSum(), red);
    s2 = block_reduce<kImgNT>(s2, OpSum(), red);
    cmin = block_reduce<kImgNT>(cmin, OpSum(), red);
    cmax = block_reduce<kImgNT>(cmax, OpSum(), red);
    const int b = blockIdx.x;
    if (linear_tid() == 0) {
        part[(0 * R + r) * nb + b] = mse; part[(1 * R + r) * nb + b] = s1; part[(2 * R + r) * nb + b] = s2;
        part[(3 * R + r) * nb + b] = cmin; part[(4 * R + r) * nb + b] = cmax;
    }
    if (!last_block_ticket(counter, (unsigned)(nb * R), &last)) return;
    for (int q = 0; q < R; ++q) {
        double v[5] = {0, 0, 0, 0, 0};
        for (int k = linear_tid(); k < nb; k += kImgNT)
            for (int j = 0; j < 5; ++j) v[j] += __ldcg(&part[(j * R + q) * nb + k]);
        for (int j = 0; j < 5; ++j) v[j] = block_reduce<kImgNT>(v[j], OpSum(), red);
        if (linear_tid() == 0) {
            stats_out[q].mse = v[0] / (double)HW;
            stats_out[q].s1 = v[1]; stats_out[q].s2 = v[2]; stats_out[q].cnt_min = v[3]; stats_out[q].cnt_max = v[4];
        }
    }
    if (linear_tid() == 0) *counter = 0;
}

// ---- C ---------------------------------------------------------------------------------------------------
// dLdI_r = coefA[r] * (corr2d(Gx,Kx) + corr2d(Gy,Ky)) + gN/D + g_m [I==min]/#min + g_M [I==max]/#max
__global__ void __launch_bounds__(kImgNT)
k_img_C(const double* __restrict__ imgs, const double* __restrict__ edges, const double* __restrict__ gNdiv, int H, int W,
        const Stats* __restrict__ stats, const double* __restrict__ coefA, const double* __restrict__ coefB,
        double* __restrict__ dldi, float* __restrict__ dldi32 /* optional: (float)(dldi / 2 pi) for the fast backward */) {
    __shared__ double tile[kImgTY + 4][kImgTX + 4];
    __shared__ double gxs[kImgTY + 2][kImgTX + 2], gys[kImgTY + 2][kImgTX + 2];
    const int r = blockIdx.z;
    const int64_t HW = (int64_t)H * W;
    const double* img = imgs + r * HW;
    const int x0 = blockIdx.x * kImgTX, y0 = blockIdx.y * kImgTY;
    load_tile<2>(img, H, W, x0, y0, tile);
    __syncthreads();
    constexpr int GW = kImgTX + 2, GH = kImgTY + 2;
    for (int k = linear_tid(); k < GW * GH; k += kImgNT) {
        const int ly = k / GW, lx = k % GW;
        const int y = y0 + ly - 1, x = x0 + lx - 1;
        double gx = 0.0, gy = 0.0;
        if (x >= 0 && x < W && y >= 0 && y < H) scharr_at(&tile[ly + 1][lx + 1], kImgTX + 4, gx, gy);
        gxs[ly][lx] = gx; gys[ly][lx] = gy;     // zero outside the image: the forward's 'same' output has no such rows
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;
    const Stats st = stats[r];
    const double I = tile[threadIdx.y + 2][threadIdx.x + 2];
    const int64_t p = (int64_t)y * W + x;
    const double adj = scharr_adjoint_at(&gxs[threadIdx.y + 1][threadIdx.x + 1], &gys[threadIdx.y + 1][threadIdx.x + 1], GW);
    const double c = I - st.mn;
    double gN = coefB[r] * (edges[r * HW + p] - c / st.D);
    if (gNdiv != nullptr) gN += gNdiv[r * HW + p];
    const double g_M = -st.s2 / (st.D * st.D);
    const double g_m = -st.s1 / st.D + st.s2 / (st.D * st.D);
    double out = coefA[r] * adj + gN / st.D;
    if (I == st.mn) out += g_m / st.cnt_min;
    if (I == st.mx) out += g_M / st.cnt_max;
    dldi[r * HW + p] = out;
    if (dldi32 != nullptr) dldi32[r * HW + p] = (float)(out * kInv2Pi);
}

// ---- D1: S = divk(Gx(N)) + divk(Gy(N)); div = mean|S|; sbar = coefD[r] * sign(S) (coefD may be NULL: forward only)
__global__ void __launch_bounds__(kImgNT)
k_img_D1(const double* __restrict__ imgs, int64_t img_stride, int H, int W, int nb, const Stats* stats_in,
         int stats_in_stride, const double* __restrict__ coefD, double* __restrict__ sbar, double* __restrict__ part,
         Stats* stats_out, unsigned int* __restrict__ counter) {
    __shared__ double tile[kImgTY + 4][kImgTX + 4];
    __shared__ double gxs[kImgTY + 2][kImgTX + 2], gys[kImgTY + 2][kImgTX + 2];
    __shared__ double red[8];
    __shared__ bool last;
    const int r = blockIdx.z, R = gridDim.z;
    const int64_t HW = (int64_t)H * W;
    const double* img = imgs + r * img_stride;
    const Stats st = stats_in[r * stats_in_stride];
    const int x0 = blockIdx.x * kImgTX, y0 = blockIdx.y * kImgTY;
    load_tile<2>(img, H, W, x0, y0, tile);
    __syncthreads();
    // normalise in place (inside the image only; the zero padding stays zero)
    constexpr int PW = kImgTX + 4, PH = kImgTY + 4;
    for (int k = linear_tid(); k < PW * PH; k += kImgNT) {
        const int ly = k / PW, lx = k % PW;
        const int y = y0 + ly - 2, x = x0 + lx - 2;
        if (x >= 0 && x < W && y >= 0 && y < H) tile[ly][lx] = (tile[ly][lx] - st.mn) / st.D;
    }
    __syncthreads();
    constexpr int GW = kImgTX + 2, GH = kImgTY + 2;
    for (int k = linear_tid(); k < GW * GH; k += kImgNT) {
        const int ly = k / GW, lx = k % GW;
        const int y = y0 + ly - 1, x = x0 + lx - 1;
        double gx = 0.0, gy = 0.0;
        if (x >= 0 && x < W && y >= 0 && y < H) scharr_at(&tile[ly + 1][lx + 1], PW, gx, gy);
        gxs[ly][lx] = gx; gys[ly][lx] = gy;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    double a = 0.0;
    if (x < W && y < H) {
        const double S = __dadd_rn(divk_at(&gxs[threadIdx.y + 1][threadIdx.x + 1], GW), divk_at(&gys[threadIdx.y + 1][threadIdx.x + 1], GW));
        a = fabs(S);
        if (sbar != nullptr) sbar[r * HW + (int64_t)y * W + x] = coefD[r] * sign_of(S);
    }
    a = block_reduce<kImgNT>(a, OpSum(), red);
    const int b = blockIdx.y * gridDim.x + blockIdx.x;
    if (linear_tid() == 0) part[r * nb + b] = a;
    if (!last_block_ticket(counter, (unsigned)(nb * R), &last)) return;
    for (int q = 0; q < R; ++q) {
        double s = 0.0;
        for (int k = linear_tid(); k < nb; k += kImgNT) s += __ldcg(&part[q * nb + k]);
        s = block_reduce<kImgNT>(s, OpSum(), red);
        if (linear_tid() == 0) stats_out[q].div = s / (double)HW;
    }
    if (linear_tid() == 0) *counter = 0;
}

// ---- D2: gNdiv = scharr_adjoint(kbar, kbar), kbar = divk(sbar)  (DIV_KERN is symmetric: adjoint == itself)
__global__ void __launch_bounds__(kImgNT)
k_img_D2(const double* __restrict__ sbar, int H, int W, double* __restrict__ gNdiv) {
    __shared__ double tile[kImgTY + 4][kImgTX + 4];
    __shared__ double kb[kImgTY + 2][kImgTX + 2];
    const int r = blockIdx.z;
    const int64_t HW = (int64_t)H * W;
    const int x0 = blockIdx.x * kImgTX, y0 = blockIdx.y * kImgTY;
    load_tile<2>(sbar + r * HW, H, W, x0, y0, tile);
    __syncthreads();
    constexpr int GW = kImgTX + 2, GH = kImgTY + 2;
    for (int k = linear_tid(); k < GW * GH; k += kImgNT) {
        const int ly = k / GW, lx = k % GW;
        const int y = y0 + ly - 1, x = x0 + lx - 1;
        kb[ly][lx] = (x >= 0 && x < W && y >= 0 && y < H) ? divk_at(&tile[ly + 1][lx + 1], kImgTX + 4) : 0.0;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;
    const double* p = &kb[threadIdx.y + 1][threadIdx.x + 1];
    gNdiv[r * HW + (int64_t)y * W + x] = scharr_adjoint_at(p, p, GW);
}

// ---- scalar epilogue: K8 of SURVEY.md §2.1 (reference src/eincm/losses.py:171-193) ------------------------
// Assembles the loss from the per-reference statistics; also derives the cotangent scales for the backward.
// mode 0: compute coefA/B/D from the zero-IWE constants (called before B / C); mode 1: final loss.
// coefA/B/D from the zero-IWE constants (needed by the image backward)
__device__ __forceinline__ void scalars_coefs(DevScalars* sc, int R, double HW, double alpha, double beta, double delta, int use_div) {
    const double C0 = sc->zero[0].contrast;
    const double D0 = sc->zero[0].div;
    for (int r = 0; r < R; ++r) {
        const double w = sc->weights[r];
        const double a_r = -alpha * w / ((C0 + kEps) * R);
        const double b_r = beta * w / ((-sc->zero[r].mse + kEps) * R);
        const double d_r = use_div ? delta * w / ((D0 + kEps) * R) : 0.0;
        sc->coefA[r] = a_r * (2.0 / HW);
        sc->coefB[r] = b_r * (-2.0 / HW);
        sc->coefD[r] = d_r / HW;
    }
}

// final loss (reference src/eincm/losses.py:171-193) from the per-reference statistics
__device__ __forceinline__ void scalars_loss(DevScalars* sc, int R, double alpha, double beta, double gamma, double delta,
                                             int use_tv, int use_div, double* loss_out) {
    const double C0 = sc->zero[0].contrast;
    const double D0 = sc->zero[0].div;
    double s_corr = 0.0, s_con = 0.0, s_div = 0.0;
    for (int r = 0; r < R; ++r) {
        const double w = sc->weights[r];
        s_corr += (w * (-sc->ref[r].mse)) / ((-sc->zero[r].mse) + kEps);         // losses.py:176
        s_con += (w * sc->ref[r].contrast) / (C0 + kEps);                        // losses.py:177
        if (use_div) s_div += (w * sc->ref[r].div) / (D0 + kEps);                // losses.py:178
    }
    const double mean_rel_corr = s_corr / R, mean_rel_contrast = s_con / R, mean_rel_div = s_div / R;
    const double tv = use_tv ? sc->tv_sum / (sc->tv_cnt + kEps) : 0.0;          // regularizers.py:31-36, losses.py:171
    const double contrast_loss = mean_rel_contrast * (-1.0), correlation_loss = mean_rel_corr * (-1.0);
    const double loss = (alpha * contrast_loss + beta * correlation_loss) + (gamma * tv + delta * mean_rel_div);
    sc->loss = loss; sc->mean_rel_corr = mean_rel_corr; sc->mean_rel_contrast = mean_rel_contrast;
    sc->mean_rel_div = mean_rel_div; sc->tv = tv;
    if (loss_out != nullptr) *loss_out = loss;
}

// mode 0: cotangent scales (called before B / C); mode 1: final loss.
__global__ void k_scalars(DevScalars* __restrict__ sc, int R, double HW, double alpha, double beta, double gamma, double delta,
                          int use_tv, int use_div, int mode, double* __restrict__ loss_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (mode == 0) scalars_coefs(sc, R, HW, alpha, beta, delta, use_div);
    else scalars_loss(sc, R, alpha, beta, gamma, delta, use_tv, use_div, loss_out);
}

// ---- gradient epilogue: reduce the gather partials (or pass the scatter buffer through), write grad_out and
// accumulate d loss / d alpha_handover = <grad, prev - theta>  (reference src/eincm/losses.py:269)
__global__ void __launch_bounds__(256)
k_grad_out(const double2* __restrict__ partial, int S, const double* __restrict__ grad_buf, int n_elems /* h*w */,
           const double* __restrict__ prev, const double* __restrict__ theta, double* __restrict__ grad_out,
           DevScalars* __restrict__ sc) {
    __shared__ double red[8];
    double da = 0.0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += gridDim.x * blockDim.x) {
        double g0, g1;
        if (partial != nullptr) {
            g0 = 0.0; g1 = 0.0;
            for (int s = 0; s < S; ++s) { const double2 v = partial[e * S + s]; g0 += v.x; g1 += v.y; }
        } else {
            g0 = grad_buf[2 * e]; g1 = grad_buf[2 * e + 1];
        }
        if (grad_out != nullptr) { grad_out[2 * e] = g0; grad_out[2 * e + 1] = g1; }
        if (prev != nullptr) da += g0 * (prev[2 * e] - theta[2 * e]) + g1 * (prev[2 * e + 1] - theta[2 * e + 1]);
    }
    if (prev != nullptr) {
        da = block_reduce<256>(da, OpSum(), red);
        if (threadIdx.x == 0) atomicAdd(&sc->dalpha, da);
    }
}

}  // namespace eincm
