// theta <-> sensor-size kernels: bilinear up-scaling of the tile field (forward) and its transpose (backward),
// plus the total-variation regulariser on the masked dense field.
//   forward : reference src/utils/theta_utils.py:10-37 (jax.image.scale_and_translate, method='bilinear')
//   backward: the einsum-transpose autodiff derives for it  (dtheta = Wy^T . G . Wx)
//   TV      : reference src/eincm/regularizers.py:14-38 + src/utils/theta_utils.py:40-73
#pragma once
#include "common.cuh"

namespace eincm {

// Per-axis resize taps.  For an output coordinate o: out[o] = w0 * in[i0] + w1 * in[i1]  (<= 2 taps when
// up-scaling; i1 == i0 and w1 == 0 when there is a single tap).  Built on the host exactly like
// jax._src.image.scale.compute_weight_mat (see eincm_plan.cu: build_axis_taps).
struct AxisTaps {
    const int* i0;      // [n_out]
    const int* i1;      // [n_out]
    const double* w0;   // [n_out]
    const double* w1;   // [n_out]
    const int* lo;      // [n_in]  first output index whose taps touch input i
    const int* hi;      // [n_in]  one past the last
};

// theta_full[y][x][c] = sum_ij Wy[i][y] Wx[j][x] theta[i][j][c]; optional handover blend
// theta_ho = a*prev + (1-a)*theta (reference src/eincm/losses.py:269) applied to the taps on the fly.
__global__ void k_upsample_theta(const double* __restrict__ theta, const double* __restrict__ prev, double a_ho,
                                 int h, int w, int H, int W, AxisTaps ty, AxisTaps tx, double2* __restrict__ theta_full) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int i0 = ty.i0[y], i1 = ty.i1[y], j0 = tx.i0[x], j1 = tx.i1[x];
    const double wy0 = ty.w0[y], wy1 = ty.w1[y], wx0 = tx.w0[x], wx1 = tx.w1[x];
    double out[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        double t00 = theta[(i0 * w + j0) * 2 + c], t01 = theta[(i0 * w + j1) * 2 + c];
        double t10 = theta[(i1 * w + j0) * 2 + c], t11 = theta[(i1 * w + j1) * 2 + c];
        if (prev != nullptr) {
            const double b = 1.0 - a_ho;
            t00 = a_ho * prev[(i0 * w + j0) * 2 + c] + b * t00;
            t01 = a_ho * prev[(i0 * w + j1) * 2 + c] + b * t01;
            t10 = a_ho * prev[(i1 * w + j0) * 2 + c] + b * t10;
            t11 = a_ho * prev[(i1 * w + j1) * 2 + c] + b * t11;
        }
        out[c] = wy0 * (wx0 * t00 + wx1 * t01) + wy1 * (wx0 * t10 + wx1 * t11);
    }
    theta_full[y * W + x] = make_double2(out[0], out[1]);
}

// ---- backward of the resize, gather form (small theta: h*w <= kGatherMaxTiles) --------------------------
// One CTA per (theta element, split s): sums Wy[i][y] Wx[j][x] (G[y][x] + tv_coef * Gtv[y][x]) over its share
// of the element's support rectangle; partial[(i*w+j)*S + s] = (sum_c0, sum_c1).  Deterministic.
constexpr int kGatherMaxTiles = 4096;

__global__ void __launch_bounds__(256)
k_theta_grad_gather(const double2* __restrict__ G, const double2* __restrict__ Gtv, const DevScalars* __restrict__ sc,
                    double gamma, int h, int w, int H, int W, int S, AxisTaps ty, AxisTaps tx, double2* __restrict__ partial) {
    __shared__ double sh[8];
    const int ij = blockIdx.x / S, s = blockIdx.x % S;
    const int i = ij / w, j = ij % w;
    const int ylo = ty.lo[i], yhi = ty.hi[i], xlo = tx.lo[j], xhi = tx.hi[j];
    const int ny = yhi - ylo;
    const int y_begin = ylo + (int)(((long long)ny * s) / S), y_end = ylo + (int)(((long long)ny * (s + 1)) / S);
    const int nx = xhi - xlo;
    double tvc = 0.0;
    if (Gtv != nullptr) tvc = gamma * 0.25 / (sc->tv_cnt + kEps);
    double a0 = 0.0, a1 = 0.0;
    const int total = (y_end - y_begin) * nx;
    for (int k = threadIdx.x; k < total; k += blockDim.x) {
        const int y = y_begin + k / nx, x = xlo + k % nx;
        const double wy = (ty.i0[y] == i ? ty.w0[y] : 0.0) + ((ty.i1[y] == i && ty.i1[y] != ty.i0[y]) ? ty.w1[y] : 0.0);
        const double wx = (tx.i0[x] == j ? tx.w0[x] : 0.0) + ((tx.i1[x] == j && tx.i1[x] != tx.i0[x]) ? tx.w1[x] : 0.0);
        double2 g = G[y * W + x];
        if (Gtv != nullptr) { const double2 t = Gtv[y * W + x]; g.x += tvc * t.x; g.y += tvc * t.y; }
        const double ww = wy * wx;
        a0 += ww * g.x;
        a1 += ww * g.y;
    }
    a0 = block_reduce<256>(a0, OpSum(), sh);
    a1 = block_reduce<256>(a1, OpSum(), sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = make_double2(a0, a1);
}

// ---- backward of the resize, scatter form (large / dense theta) -----------------------------------------
// One thread per sensor pixel adds its <= 4 weighted contributions into grad_buf[h][w][2] (pre-zeroed).
__global__ void k_theta_grad_scatter(const double2* __restrict__ G, const double2* __restrict__ Gtv, const DevScalars* __restrict__ sc,
                                     double gamma, int h, int w, int H, int W, AxisTaps ty, AxisTaps tx, double* __restrict__ grad_buf) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    double2 g = G[y * W + x];
    if (Gtv != nullptr) {
        const double tvc = gamma * 0.25 / (sc->tv_cnt + kEps);
        const double2 t = Gtv[y * W + x];
        g.x += tvc * t.x; g.y += tvc * t.y;
    }
    if (g.x == 0.0 && g.y == 0.0) return;
    const int i0 = ty.i0[y], i1 = ty.i1[y], j0 = tx.i0[x], j1 = tx.i1[x];
    const double wy0 = ty.w0[y], wy1 = ty.w1[y], wx0 = tx.w0[x], wx1 = tx.w1[x];
    const bool single_y = (i1 == i0) || wy1 == 0.0, single_x = (j1 == j0) || wx1 == 0.0;
    if (single_y && single_x && wy0 == 1.0 && wx0 == 1.0 && h == H && w == W) {   // dense theta: identity resize
        grad_buf[(i0 * w + j0) * 2 + 0] = g.x;
        grad_buf[(i0 * w + j0) * 2 + 1] = g.y;
        return;
    }
    atomicAdd(&grad_buf[(i0 * w + j0) * 2 + 0], wy0 * wx0 * g.x);
    atomicAdd(&grad_buf[(i0 * w + j0) * 2 + 1], wy0 * wx0 * g.y);
    if (!single_x) {
        atomicAdd(&grad_buf[(i0 * w + j1) * 2 + 0], wy0 * wx1 * g.x);
        atomicAdd(&grad_buf[(i0 * w + j1) * 2 + 1], wy0 * wx1 * g.y);
    }
    if (!single_y) {
        atomicAdd(&grad_buf[(i1 * w + j0) * 2 + 0], wy1 * wx0 * g.x);
        atomicAdd(&grad_buf[(i1 * w + j0) * 2 + 1], wy1 * wx0 * g.y);
        if (!single_x) {
            atomicAdd(&grad_buf[(i1 * w + j1) * 2 + 0], wy1 * wx1 * g.x);
            atomicAdd(&grad_buf[(i1 * w + j1) * 2 + 1], wy1 * wx1 * g.y);
        }
    }
}

// ---- total variation of the masked flow (only when gamma != 0 and cur_pyr_lvl <= 0) ---------------------
// flow = theta_full * mask; a,b = Scharr(flow_x); c,d = Scharr(flow_y)
// tv_sum = sum 0.25(|a|+|b|) + 0.25(|c|+|d|); tv_cnt = #pixels with any non-zero gradient
// Gtv[p] = (adjoint(sign a, sign b), adjoint(sign c, sign d)) * mask   (scaled by gamma*0.25/(cnt+eps) later)
constexpr int kTvTX = 32, kTvTY = 8;

__global__ void __launch_bounds__(kTvTX * kTvTY)
k_tv(const double2* __restrict__ theta_full, const uint8_t* __restrict__ mask, int H, int W, int n_blocks,
     double2* __restrict__ Gtv, double* __restrict__ part /* [2][n_blocks] */, DevScalars* __restrict__ sc) {
    constexpr int PW = kTvTX + 4, PH = kTvTY + 4;       // flow tile with halo 2
    constexpr int GW = kTvTX + 2, GH = kTvTY + 2;       // sign tiles with halo 1
    __shared__ double fx[PH][PW], fy[PH][PW];
    __shared__ double sa[GH][GW], sb[GH][GW], sc_[GH][GW], sd[GH][GW];
    __shared__ double red[8];
    __shared__ bool is_last;
    const int tid = linear_tid();
    const int x0 = blockIdx.x * kTvTX, y0 = blockIdx.y * kTvTY;
    for (int k = tid; k < PW * PH; k += kTvTX * kTvTY) {
        const int ly = k / PW, lx = k % PW;
        const int y = y0 + ly - 2, x = x0 + lx - 2;
        double vx = 0.0, vy = 0.0;
        if (x >= 0 && x < W && y >= 0 && y < H && mask[y * W + x]) { const double2 t = theta_full[y * W + x]; vx = t.x; vy = t.y; }
        fx[ly][lx] = vx; fy[ly][lx] = vy;
    }
    __syncthreads();
    double my_sum = 0.0, my_cnt = 0.0;
    for (int k = tid; k < GW * GH; k += kTvTX * kTvTY) {
        const int ly = k / GW, lx = k % GW;
        const int y = y0 + ly - 1, x = x0 + lx - 1;
        double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
        if (x >= 0 && x < W && y >= 0 && y < H) {        // gradients exist only inside the image ('same' output)
            scharr_at(&fx[ly + 1][lx + 1], PW, a, b);
            scharr_at(&fy[ly + 1][lx + 1], PW, c, d);
            if (ly >= 1 && ly <= kTvTY && lx >= 1 && lx <= kTvTX) {   // own pixel of this tile
                my_sum += __dadd_rn(__dadd_rn(__dmul_rn(fabs(a), 0.25), __dmul_rn(fabs(b), 0.25)),
                                    __dadd_rn(__dmul_rn(fabs(c), 0.25), __dmul_rn(fabs(d), 0.25)));
                my_cnt += (fabs(a) > 0.0 || fabs(b) > 0.0 || fabs(c) > 0.0 || fabs(d) > 0.0) ? 1.0 : 0.0;
            }
        }
        sa[ly][lx] = sign_of(a); sb[ly][lx] = sign_of(b); sc_[ly][lx] = sign_of(c); sd[ly][lx] = sign_of(d);
    }
    __syncthreads();
    {
        const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
        if (x < W && y < H) {
            double gx = 0.0, gy = 0.0;
            if (mask[y * W + x]) {
                gx = scharr_adjoint_at(&sa[threadIdx.y + 1][threadIdx.x + 1], &sb[threadIdx.y + 1][threadIdx.x + 1], GW);
                gy = scharr_adjoint_at(&sc_[threadIdx.y + 1][threadIdx.x + 1], &sd[threadIdx.y + 1][threadIdx.x + 1], GW);
            }
            Gtv[y * W + x] = make_double2(gx, gy);
        }
    }
    const int b = blockIdx.y * gridDim.x + blockIdx.x;
    my_sum = block_reduce<kTvTX * kTvTY>(my_sum, OpSum(), red);
    my_cnt = block_reduce<kTvTX * kTvTY>(my_cnt, OpSum(), red);
    if (tid == 0) {
        part[b] = my_sum; part[n_blocks + b] = my_cnt;
        __threadfence();
        is_last = (atomicAdd(&sc->counters[3], 1u) == (unsigned)(n_blocks - 1));
    }
    __syncthreads();
    if (is_last) {
        double s = 0.0, c = 0.0;
        for (int k = tid; k < n_blocks; k += kTvTX * kTvTY) { s += __ldcg(&part[k]); c += __ldcg(&part[n_blocks + k]); }
        s = block_reduce<kTvTX * kTvTY>(s, OpSum(), red);
        c = block_reduce<kTvTX * kTvTY>(c, OpSum(), red);
        if (tid == 0) { sc->tv_sum = s; sc->tv_cnt = c; sc->counters[3] = 0; }
    }
}

}  // namespace eincm
