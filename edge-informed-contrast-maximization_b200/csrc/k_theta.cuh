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

// The flow operand of one evaluation: tile field theta [h][w][2] (null = zero flow), optional handover blend
// theta_ho = a*prev + (1-a)*theta (reference src/eincm/losses.py:269), and the per-axis resize taps to the sensor size.
struct ThetaSrc {
    const double* theta;
    const double* prev;
    double a_ho;
    int h, w;
    AxisTaps ty, tx;
};

// theta_full[y][x][c] = sum_ij Wy[i][y] Wx[j][x] theta[i][j][c] (reference src/utils/theta_utils.py:25-35).  The one
// definition used by k_upsample_theta (dense field for the TV regulariser and the debug tap) and by the event kernels,
// which evaluate it per source tile instead of reading a dense field: identical arithmetic, identical bits.
__device__ __forceinline__ double2 theta_at(const ThetaSrc& T, int x, int y) {
    const int i0 = __ldg(T.ty.i0 + y), i1 = __ldg(T.ty.i1 + y), j0 = __ldg(T.tx.i0 + x), j1 = __ldg(T.tx.i1 + x);
    const double wy0 = __ldg(T.ty.w0 + y), wy1 = __ldg(T.ty.w1 + y), wx0 = __ldg(T.tx.w0 + x), wx1 = __ldg(T.tx.w1 + x);
    const int w = T.w;
    // scalar loads: the C-ABI only promises 8-byte alignment of theta / prev
    auto ld2 = [](const double* p, int e) { return make_double2(__ldg(p + 2 * e), __ldg(p + 2 * e + 1)); };
    double2 t00 = ld2(T.theta, i0 * w + j0), t01 = ld2(T.theta, i0 * w + j1), t10 = ld2(T.theta, i1 * w + j0), t11 = ld2(T.theta, i1 * w + j1);
    if (T.prev != nullptr) {
        const double a = T.a_ho, b = 1.0 - T.a_ho;
        const double2 p00 = ld2(T.prev, i0 * w + j0), p01 = ld2(T.prev, i0 * w + j1), p10 = ld2(T.prev, i1 * w + j0), p11 = ld2(T.prev, i1 * w + j1);
        t00.x = a * p00.x + b * t00.x; t00.y = a * p00.y + b * t00.y;
        t01.x = a * p01.x + b * t01.x; t01.y = a * p01.y + b * t01.y;
        t10.x = a * p10.x + b * t10.x; t10.y = a * p10.y + b * t10.y;
        t11.x = a * p11.x + b * t11.x; t11.y = a * p11.y + b * t11.y;
    }
    double2 out;
    out.x = wy0 * (wx0 * t00.x + wx1 * t01.x) + wy1 * (wx0 * t10.x + wx1 * t11.x);
    out.y = wy0 * (wx0 * t00.y + wx1 * t01.y) + wy1 * (wx0 * t10.y + wx1 * t11.y);
    return out;
}

__global__ void k_upsample_theta(const ThetaSrc T, int H, int W, double2* __restrict__ theta_full) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    theta_full[y * W + x] = theta_at(T, x, y);
}

// ---- backward of the resize, gather form (tile theta: h*w <= kGatherMaxTiles) ------------------------------------------
// d theta[i][j] = sum_{y,x} Wy[i][y] Wx[j][x] (G[y][x] + tv_coef * Gtv[y][x]).  One WARP per work item (theta element, row
// split sy, column split sx): <= 32 rows (one row weight per lane, broadcast by shuffle) x <= 96 columns (three per lane), no
// block-level synchronisation; the partial goes to grad[i][j] (pre-zeroed) with two float64 reductions.  The last warp to
// finish evaluates d loss / d alpha_handover = <grad, prev - theta> (reference src/eincm/losses.py:269) from the complete
// gradient, so the whole theta backward is one launch.
constexpr int kGatherMaxTiles = 4096;
constexpr int kTgWarps = 4;              // warps (work items) per CTA
#ifndef EINCM_TG_TRIPS
#define EINCM_TG_TRIPS 2
#endif
constexpr int kTgTrips = EINCM_TG_TRIPS;   // row groups a work item walks through, one after the other (16x16 theta on 640x480: 1 -> 14.9 us in two
                                           // waves of CTAs, 2 -> 12.8 us in one, 4 -> 14.6 us)
constexpr int kTgRows = 4, kTgCols = 96;  // rows x columns per work item: 12 independent 16-byte loads in flight per lane

__device__ __forceinline__ double axis_weight(const AxisTaps& t, int o, int i) {
    const int i0 = __ldg(t.i0 + o), i1 = __ldg(t.i1 + o);
    const double w0 = __ldg(t.w0 + o), w1 = __ldg(t.w1 + o);
    return (i0 == i ? w0 : 0.0) + ((i1 == i && i1 != i0) ? w1 : 0.0);
}

template <bool HAS_TV>
__device__ __forceinline__ void theta_grad_body(const double2* __restrict__ G, const double2* __restrict__ Gtv, DevScalars* __restrict__ sc,
             double gamma, int h, int w, int H, int W, int SY, int SX, int n_items, const AxisTaps& ty, const AxisTaps& tx,
             const double* __restrict__ prev, const double* __restrict__ theta, double* __restrict__ grad /* [h][w][2] */,
             const double* __restrict__ loss_dev, double* __restrict__ host_out /* mapped pinned memory or null */, int host_grad,
             const int* __restrict__ skip = nullptr) {
    // programmatic dependent launch (no-ops when launched plainly)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (skip != nullptr && *reinterpret_cast<const volatile int*>(skip) != 0) return;
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * kTgWarps + (threadIdx.x >> 5);
    if (item >= n_items) return;
    const int ij = item / (SY * SX), rem = item - ij * (SY * SX);
    const int sy = rem / SX, sx = rem - sy * SX;
    const int i = ij / w, j = ij - i * w;
    const int ylo = __ldg(ty.lo + i), yhi = __ldg(ty.hi + i), xlo = __ldg(tx.lo + j), xhi = __ldg(tx.hi + j);
    const int ny = yhi - ylo, nx = xhi - xlo;
    const int y_begin = ylo + (ny * sy) / SY, y_end = ylo + (ny * (sy + 1)) / SY;
    const int x_begin = xlo + (nx * sx) / SX, x_end = xlo + (nx * (sx + 1)) / SX;
    const double tvc = HAS_TV ? gamma * 0.25 / (sc->tv_cnt + kEps) : 0.0;
    double a0 = 0.0, a1 = 0.0;
    for (int yb = y_begin; yb < y_end; yb += kTgRows) {                 // one trip by construction (host sizes SY); general anyway
        for (int xb = x_begin; xb < x_end; xb += kTgCols) {
            // all loads first: kTgRows x 3 cells per lane (out-of-range cells read an in-range address with weight 0)
            double2 g[kTgRows][3], t[kTgRows][3];
            int xc[3];
            bool okc[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) { xc[c] = xb + lane + 32 * c; okc[c] = xc[c] < x_end; if (!okc[c]) xc[c] = x_begin; }
#pragma unroll
            for (int r = 0; r < kTgRows; ++r) {
                const int y = min(yb + r, y_end - 1);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    g[r][c] = __ldcg(G + (int64_t)y * W + xc[c]);
                    if (HAS_TV) t[r][c] = __ldcg(Gtv + (int64_t)y * W + xc[c]);
                }
            }
            double wy[kTgRows], wx[3];
#pragma unroll
            for (int r = 0; r < kTgRows; ++r) wy[r] = (yb + r < y_end) ? axis_weight(ty, yb + r, i) : 0.0;
#pragma unroll
            for (int c = 0; c < 3; ++c) wx[c] = okc[c] ? axis_weight(tx, xc[c], j) : 0.0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int r = 0; r < kTgRows; ++r) {
                    double gx = g[r][c].x, gy = g[r][c].y;
                    if (HAS_TV) { gx += tvc * t[r][c].x; gy += tvc * t[r][c].y; }
                    c0 += wy[r] * gx;
                    c1 += wy[r] * gy;
                }
                a0 += wx[c] * c0;
                a1 += wx[c] * c1;
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
    unsigned int ticket = 0u;
    if (lane == 0) {
        atomicAdd(grad + 2 * ij, a0);
        atomicAdd(grad + 2 * ij + 1, a1);
        __threadfence();
        ticket = atomicAdd(&sc->counters[5], 1u);
    }
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket != (unsigned int)n_items - 1u) return;
    double da = 0.0;
    if (prev != nullptr) {
        const int n = h * w * 2;
        for (int e = lane; e < n; e += 32) da += __ldcg(grad + e) * (prev[e] - theta[e]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) da += __shfl_xor_sync(0xffffffffu, da, o);
    if (lane == 0) { sc->dalpha = da; sc->counters[5] = 0u; }
    // synchronous host entry points: the last warp delivers [gradient | loss | d/d alpha] straight into mapped pinned host
    // memory (no device -> host copy operation after the kernel)
    if (host_out != nullptr) {                  // layout: [sequence | loss | d alpha | gradient ...] (as k_backward_fold)
        const int n = host_grad ? h * w * 2 : 0;
        for (int e = lane; e < n; e += 32) host_out[3 + e] = __ldcg(grad + e);
        if (lane == 0) { host_out[1] = __ldcg(loss_dev); host_out[2] = da; }
        __threadfence_system();
        __syncwarp();
        // sequence number of this evaluation, written after everything else is visible to the host: the host entry point
        // polls it in its own memory instead of calling into the driver (no lock shared with threads that are launching)
        if (lane == 0) {
            const double seq = sc->eval_seq + 1.0;
            sc->eval_seq = seq;
            *reinterpret_cast<volatile double*>(host_out) = seq;
            __threadfence_system();
        }
    }
}

template <bool HAS_TV>
__global__ void __launch_bounds__(kTgWarps * 32)
k_theta_grad(const double2* __restrict__ G, const double2* __restrict__ Gtv, DevScalars* __restrict__ sc,
             double gamma, int h, int w, int H, int W, int SY, int SX, int n_items, AxisTaps ty, AxisTaps tx,
             const double* __restrict__ prev, const double* __restrict__ theta, double* __restrict__ grad /* [h][w][2] */,
             const double* __restrict__ loss_dev, double* __restrict__ host_out /* mapped pinned memory or null */, int host_grad,
             const int* __restrict__ skip /* non-zero: return at once (or null) */) {
    theta_grad_body<HAS_TV>(G, Gtv, sc, gamma, h, w, H, W, SY, SX, n_items, ty, tx, prev, theta, grad, loss_dev, host_out, host_grad, skip);
}

// batched form (blockIdx.y = window; gamma == 0): one argument record per window in device memory
struct ThetaGradArgs {
    const double2* G; DevScalars* sc; int h, w, H, W, SY, SX, n_items, host_grad; AxisTaps ty, tx;
    const double* prev; const double* theta; double* grad; const double* loss_dev; double* host_out;
    const int* skip;                // non-zero: the window's CTAs return at once (batched solve graphs), or null
};

// Window of a CTA of the batched kernels: blockIdx.y, or - when a batched solve graph hands over the list of windows whose level has not
// ended (order[0] = their number, order[1 ..] = their indices: k_bfgs_compact) - the blockIdx.y-th of those; -1: no window, return at once.
__device__ __forceinline__ int batch_window(const int* __restrict__ order) {
    int win = (int)blockIdx.y;
    if (order != nullptr) {                          // written by an earlier kernel of the same graph launch: read through L2, never the read-only path
        if (win >= __ldcg(order)) return -1;
        win = __ldcg(order + 1 + win);
    }
    return win;
}

__global__ void __launch_bounds__(kTgWarps * 32)
k_theta_grad_b(const ThetaGradArgs* __restrict__ args, const int* __restrict__ order) {
    __shared__ ThetaGradArgs sA;
    const int win = batch_window(order);
    if (win < 0) return;
    {   // (load_args of k_events_tile.cuh: this header comes first)
        const uint32_t* s = reinterpret_cast<const uint32_t*>(args + win);
        uint32_t* d = reinterpret_cast<uint32_t*>(&sA);
        for (int i = threadIdx.x; i < (int)(sizeof(ThetaGradArgs) / 4); i += blockDim.x) d[i] = __ldg(s + i);
        __syncthreads();
    }
    theta_grad_body<false>(sA.G, nullptr, sA.sc, 0.0, sA.h, sA.w, sA.H, sA.W, sA.SY, sA.SX, sA.n_items, sA.ty, sA.tx, sA.prev, sA.theta, sA.grad,
                           sA.loss_dev, sA.host_out, sA.host_grad, sA.skip);
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
