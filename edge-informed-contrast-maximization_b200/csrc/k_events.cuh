// Event-space kernels: warp every event to all R reference times in registers, vote its 3x3 Gaussian-PDF
// patch into the R images of warped events (forward), and gather d loss / d IWE back through the same taps into
// the dense flow-field gradient (backward).
//   warp   : reference src/eincm/event_warpers.py:28-35      x' = x - theta_full[y,x,0] * (t - t_ref) * 1.0
//   splat  : reference src/utils/event_utils.py:31-61        9 x frame.at[r,c].add(pdf, mode='drop')
//   vmap over reference times: reference src/eincm/losses.py:26-27,58,61
// Coordinates and rint() are float64 so the event->pixel index stream is bit-exact with the float64 reference.
#pragma once
#include "common.cuh"

namespace eincm {

struct Warped { double xw, yw; int rx, ry; bool ok; };

// event_warpers.py:34-35 followed by event_utils.py:33 (jnp.round = half-to-even -> cvt.rni)
__device__ __forceinline__ Warped warp_event(int x, int y, double thx, double thy, double dt) {
    Warped o;
    o.xw = __dsub_rn((double)x, __dmul_rn(thx, dt));   // (theta*dt)*1.0 == theta*dt exactly (delta_time = 1.0)
    o.yw = __dsub_rn((double)y, __dmul_rn(thy, dt));
    // NaN / inf / absurdly far warps: every tap is out of range under either index rule -> dropped
    o.ok = (fabs(o.xw) < 1.0e9) && (fabs(o.yw) < 1.0e9);
    o.rx = o.ok ? __double2int_rn(o.xw) : 0;
    o.ry = o.ok ? __double2int_rn(o.yw) : 0;
    return o;
}

// The 3 per-axis factors exp(-0.5*q^2) for q = (r + d) - x', d = -1, 0, 1.  The 2-D tap value of
// event_utils.py:55-56 is separable: exp(-0.5(qx^2+qy^2) - log 2pi) = ex[dx] * ey[dy] / (2 pi).
__device__ __forceinline__ void axis_taps(int r, double xw, double q[3], double e[3]) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        q[d] = (double)(r + d - 1) - xw;
        e[d] = exp(-0.5 * q[d] * q[d]);
    }
}

// ---- forward: K2 + K3 of SURVEY.md §2.1 in one pass over the pixel-sorted event SoA ----------------------
template <bool WRAP>
__global__ void __launch_bounds__(256)
k_splat(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, int64_t n,
        const double2* __restrict__ theta_full, int H, int W, int R, RefTimes tref, double* __restrict__ iwe) {
    const int64_t HW = (int64_t)H * W;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t xy = ev_xy[e];
        if (xy == 0xffffffffu) continue;                     // padding sentinel of the sorted stream (k_prep.cuh)
        const double t = ev_t[e];
        const int x = xy & 0xffffu, y = xy >> 16;
        const double2 th = theta_full != nullptr ? theta_full[y * W + x] : make_double2(0.0, 0.0);
        for (int r = 0; r < R; ++r) {
            const Warped wp = warp_event(x, y, th.x, th.y, t - tref.t[r]);
            if (!wp.ok) continue;
            double qx[3], ex[3], qy[3], ey[3];
            axis_taps(wp.rx, wp.xw, qx, ex);
            axis_taps(wp.ry, wp.yw, qy, ey);
            double* img = iwe + r * HW;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    int rr = wp.ry + dy - 1, cc = wp.rx + dx - 1;
                    if (drop_index<WRAP>(rr, cc, H, W)) atomicAdd(&img[rr * W + cc], ex[dx] * ey[dy] * kInv2Pi);
                }
            }
        }
    }
}

// ---- backward: K10 + K11 ---------------------------------------------------------------------------------
// dL/dx'_{k,r} = sum_taps dLdI_r[rho,c] * v * (c - x')   (d/dx' of exp(-0.5 (c-x')^2) = v (c - x'); rint has no gradient)
// G[y_k, x_k, 0] += -(t_k - t_ref_r) * dL/dx'_{k,r}      (same for y), summed over r in registers, then a
// warp-segmented sum over runs of equal source pixel (events are pixel-sorted) and one RED per run.
template <bool WRAP>
__global__ void __launch_bounds__(256)
k_backward_events(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, int64_t n,
                  const double2* __restrict__ theta_full, int H, int W, int R, RefTimes tref,
                  const double* __restrict__ dldi, double* __restrict__ G /* [H][W][2] */) {
    const int64_t HW = (int64_t)H * W;
    const int lane = threadIdx.x & 31;
    const int64_t n_round = (n + 31) / 32 * 32;     // keep whole warps in the loop for the shuffles
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_round; e += (int64_t)gridDim.x * blockDim.x) {
        double gx_acc = 0.0, gy_acc = 0.0;
        uint32_t xy = 0xffffffffu;
        if (e < n) xy = ev_xy[e];
        if (xy != 0xffffffffu) {
            const double t = ev_t[e];
            const int x = xy & 0xffffu, y = xy >> 16;
            const double2 th = theta_full[y * W + x];
            for (int r = 0; r < R; ++r) {
                const double dt = t - tref.t[r];
                const Warped wp = warp_event(x, y, th.x, th.y, dt);
                if (!wp.ok) continue;
                double qx[3], ex[3], qy[3], ey[3];
                axis_taps(wp.rx, wp.xw, qx, ex);
                axis_taps(wp.ry, wp.yw, qy, ey);
                const double* img = dldi + r * HW;
                double gx = 0.0, gy = 0.0;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        int rr = wp.ry + dy - 1, cc = wp.rx + dx - 1;
                        if (drop_index<WRAP>(rr, cc, H, W)) {
                            const double g = img[rr * W + cc] * (ex[dx] * ey[dy] * kInv2Pi);
                            gx += g * qx[dx];
                            gy += g * qy[dy];
                        }
                    }
                }
                gx_acc -= dt * gx;
                gy_acc -= dt * gy;
            }
        }
        // segmented (by source pixel) inclusive suffix sum inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double ox = __shfl_down_sync(0xffffffffu, gx_acc, o);
            const double oy = __shfl_down_sync(0xffffffffu, gy_acc, o);
            const uint32_t oxy = __shfl_down_sync(0xffffffffu, xy, o);
            // the run [lane, lane+o] is uniform iff its end has the same key (keys are sorted)
            if (lane + o < 32 && oxy == xy) { gx_acc += ox; gy_acc += oy; }
        }
        const uint32_t prev = __shfl_up_sync(0xffffffffu, xy, 1);
        const bool head = (lane == 0) || (prev != xy);
        if (head && xy != 0xffffffffu) {
            const int x = xy & 0xffffu, y = xy >> 16;
            atomicAdd(&G[(y * W + x) * 2 + 0], gx_acc);
            atomicAdd(&G[(y * W + x) * 2 + 1], gy_acc);
        }
    }
}

}  // namespace eincm
