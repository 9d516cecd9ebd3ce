// Fast event-space kernels ("9-moment" splat).
//
// The 3x3 Gaussian-PDF patch of one warped event (reference src/utils/event_utils.py:41-59) is centred on the rounded
// pixel p = rint(x'), with sub-pixel offset f = x' - p in [-0.5, 0.5]^2.  Tap (i, j) in {-1,0,1}^2 has the value
//     v_ij = exp(-0.5 ((i - fx)^2 + (j - fy)^2)) / (2 pi)
//          = (1 / 2 pi) * G_i G_j * [ A * Bx^i * By^j ],    G_0 = 1, G_+-1 = exp(-0.5),
//            A = exp(-0.5 (fx^2 + fy^2)),  Bx = exp(fx),  By = exp(fy).
// So instead of nine scatter-adds to nine different pixels (three image rows), an event adds its nine "moments"
// A Bx^i By^j to ONE 48-byte record at pixel p - two 128-bit vector reductions and one scalar one (RED.E.ADD.F32x4
// on sm_100a), all inside two 32-byte sectors - and a stencil pass (k_compose9) turns the moment image into the image
// of warped events:   IWE[q] = (1/2pi) sum_ij G_i G_j C_ij[q - (i, j)].
// Coordinates, the warp and rint() stay float64 (bit-exact event->pixel indices); the sub-pixel offset and the moments are
// float32 (relative error ~2e-7, far inside the 1e-5 objective tolerance).  Events whose centre is not an interior
// pixel (border, or warped outside the sensor) take the slow path: nine scalar reductions with the reference's
// wrap/drop index rule, added to channel (0,0) of the destination record (G_0 G_0 = 1).
#pragma once
#include "common.cuh"
#include "k_events.cuh"

namespace eincm {

constexpr int kRec = 12;                       // floats per record: 9 moments + 3 pad (48 B, 16-byte aligned)
constexpr float kG1 = 0.60653065971263342f;    // exp(-0.5)
constexpr float kTwoPi = 6.2831853071795865f;

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void red_add_f32(float* addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

// moments m[(j+1)*3 + (i+1)] = A * Bx^i * By^j for the sub-pixel offset (fx, fy)
__device__ __forceinline__ void moments9(float fx, float fy, float m[9]) {
    const float A = __expf(-0.5f * (fx * fx + fy * fy));
    const float bx = __expf(fx), ibx = __expf(-fx), by = __expf(fy), iby = __expf(-fy);
    const float r0 = A * iby, r2 = A * by;
    m[0] = r0 * ibx; m[1] = r0; m[2] = r0 * bx;
    m[3] = A * ibx;  m[4] = A;  m[5] = A * bx;
    m[6] = r2 * ibx; m[7] = r2; m[8] = r2 * bx;
}

// ---- forward: warp + moment splat ------------------------------------------------------------------------------
// theta_full == nullptr: zero flow (the un-warped image of events, losses.py:54).
template <bool WRAP>
__device__ __forceinline__ void splat9_event(int x, int y, double t, double2 th, int H, int W, int R, const RefTimes& tref,
                                             float* __restrict__ C, int64_t HW) {
    for (int r = 0; r < R; ++r) {
        const Warped wp = warp_event(x, y, th.x, th.y, t - tref.t[r]);
        if (!wp.ok) continue;
        const float fx = (float)(wp.xw - (double)wp.rx), fy = (float)(wp.yw - (double)wp.ry);
        float* Cr = C + (int64_t)r * HW * kRec;
        if (wp.rx >= 1 && wp.rx <= W - 2 && wp.ry >= 1 && wp.ry <= H - 2) {
            float m[9];
            moments9(fx, fy, m);
            float* rec = Cr + ((int64_t)wp.ry * W + wp.rx) * kRec;
            red_add_v4(rec, m[0], m[1], m[2], m[3]);
            red_add_v4(rec + 4, m[4], m[5], m[6], m[7]);
            red_add_f32(rec + 8, m[8]);
        } else {
            // slow path: the reference's per-tap index rule; value 2 pi v_ij goes to channel (0,0) of the destination
#pragma unroll
            for (int i = -1; i <= 1; ++i) {
#pragma unroll
                for (int j = -1; j <= 1; ++j) {
                    int rr = wp.ry + j, cc = wp.rx + i;
                    if (drop_index<WRAP>(rr, cc, H, W)) {
                        const float qx = (float)i - fx, qy = (float)j - fy;
                        red_add_f32(Cr + ((int64_t)rr * W + cc) * kRec + 4, __expf(-0.5f * (qx * qx + qy * qy)));
                    }
                }
            }
        }
    }
}

// Two events per thread and iteration: both events' loads (packed coordinates, timestamp, then the dependent flow gather)
// are issued before any arithmetic, doubling the memory-level parallelism of a loop that is otherwise latency bound.
template <bool WRAP>
__global__ void __launch_bounds__(256)
k_splat9(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, int64_t n,
         const double2* __restrict__ theta_full, int H, int W, int R, const __grid_constant__ RefTimes tref,
         float* __restrict__ C /* [R][H*W][kRec] */) {
    const int64_t HW = (int64_t)H * W;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e0 < n; e0 += 2 * stride) {
        const int64_t e1 = e0 + stride;
        const bool has1 = e1 < n;
        const uint32_t xy0 = ev_xy[e0];
        const uint32_t xy1 = has1 ? ev_xy[e1] : 0u;
        const double t0 = ev_t[e0];
        const double t1 = has1 ? ev_t[e1] : 0.0;
        const int x0 = xy0 & 0xffffu, y0 = xy0 >> 16, x1 = xy1 & 0xffffu, y1 = xy1 >> 16;
        double2 th0 = make_double2(0.0, 0.0), th1 = make_double2(0.0, 0.0);
        if (theta_full != nullptr) {
            th0 = theta_full[y0 * W + x0];
            if (has1) th1 = theta_full[y1 * W + x1];
        }
        splat9_event<WRAP>(x0, y0, t0, th0, H, W, R, tref, C, HW);
        if (has1) splat9_event<WRAP>(x1, y1, t1, th1, H, W, R, tref, C, HW);
    }
}

// ---- compose: moment image -> image of warped events (float64) ------------------------------------------------------
// IWE[q] = (1/2pi) sum_ij G_i G_j C_ij[q - (i, j)]   (records outside the image contribute nothing)
constexpr int kCmpTX = 32, kCmpTY = 8;

__global__ void __launch_bounds__(kCmpTX * kCmpTY)
k_compose9(const float* __restrict__ C, int H, int W, double* __restrict__ iwe /* [R][H][W] */) {
    __shared__ float tile[kCmpTY + 2][kCmpTX + 2][9];
    const int r = blockIdx.z;
    const int64_t HW = (int64_t)H * W;
    const float* Cr = C + (int64_t)r * HW * kRec;
    const int x0 = blockIdx.x * kCmpTX, y0 = blockIdx.y * kCmpTY;
    const int tid = threadIdx.y * kCmpTX + threadIdx.x;
    constexpr int PW = kCmpTX + 2, PH = kCmpTY + 2;
    // each record is 3 float4; one thread moves one float4
    for (int k = tid; k < PW * PH * 3; k += kCmpTX * kCmpTY) {
        const int part = k % 3, cell = k / 3;
        const int ly = cell / PW, lx = cell % PW;
        const int yy = y0 + ly - 1, xx = x0 + lx - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (xx >= 0 && xx < W && yy >= 0 && yy < H) v = __ldcg(reinterpret_cast<const float4*>(Cr + ((int64_t)yy * W + xx) * kRec) + part);
        float* dst = &tile[ly][lx][0];
        if (part == 0) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w; }
        else if (part == 1) { dst[4] = v.x; dst[5] = v.y; dst[6] = v.z; dst[7] = v.w; }
        else { dst[8] = v.x; }
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;
    const int ly = threadIdx.y + 1, lx = threadIdx.x + 1;
    // tap (i, j) of a record centred at q - (i, j) lands on q
    double corners = 0.0, edges = 0.0;
    corners += (double)tile[ly + 1][lx + 1][0];      // centre at (x+1, y+1), tap (i=-1, j=-1)
    corners += (double)tile[ly + 1][lx - 1][2];      // centre (x-1, y+1), tap (+1, -1)
    corners += (double)tile[ly - 1][lx + 1][6];      // centre (x+1, y-1), tap (-1, +1)
    corners += (double)tile[ly - 1][lx - 1][8];      // centre (x-1, y-1), tap (+1, +1)
    edges += (double)tile[ly + 1][lx][1];            // centre (x, y+1), tap (0, -1)
    edges += (double)tile[ly][lx + 1][3];            // centre (x+1, y), tap (-1, 0)
    edges += (double)tile[ly][lx - 1][5];            // centre (x-1, y), tap (+1, 0)
    edges += (double)tile[ly - 1][lx][7];            // centre (x, y-1), tap (0, +1)
    const double centre = (double)tile[ly][lx][4];
    constexpr double g1 = 0.60653065971263342, g2 = 0.36787944117144233;   // exp(-0.5), exp(-1)
    iwe[(int64_t)r * HW + (int64_t)y * W + x] = (centre + g1 * edges + g2 * corners) * kInv2Pi;
}

// ---- backward: gather d loss / d IWE (float32 copy) through the 3x3 taps --------------------------------------------
// dL/dx' = sum_ij D[p + (i,j)] v_ij (i - fx),  v_ij = (1/2pi) G_i G_j A Bx^i By^j ;  same for y with (j - fy).
// Separable evaluation: with wx_i = G_i Bx^i, wy_j = G_j By^j,
//     s_j = sum_i D_ij wx_i,  sx_j = sum_i D_ij wx_i (i - fx)   =>   gx = A sum_j wy_j sx_j,  gy = A sum_j wy_j (j - fy) s_j.
// Slow path (non-interior centre): float64 D with the wrap/drop rule, as in k_backward_events.
template <bool WRAP>
__global__ void __launch_bounds__(256)
k_backward9(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, int64_t n,
            const double2* __restrict__ theta_full, int H, int W, int R, const __grid_constant__ RefTimes tref,
            const float* __restrict__ dldi32 /* [R][H][W], already scaled by 1/(2 pi) */, const double* __restrict__ dldi,
            double* __restrict__ G /* [H][W][2] */) {
    const int64_t HW = (int64_t)H * W;
    const int lane = threadIdx.x & 31;
    const int64_t n_round = (n + 31) / 32 * 32;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_round; e += (int64_t)gridDim.x * blockDim.x) {
        float gx_acc = 0.f, gy_acc = 0.f;
        uint32_t xy = 0xffffffffu;
        if (e < n) {
            xy = ev_xy[e];
            const double t = ev_t[e];
            const int x = xy & 0xffffu, y = xy >> 16;
            const double2 th = theta_full[y * W + x];
            for (int r = 0; r < R; ++r) {
                const double dt = t - tref.t[r];
                const Warped wp = warp_event(x, y, th.x, th.y, dt);
                if (!wp.ok) continue;
                const float fx = (float)(wp.xw - (double)wp.rx), fy = (float)(wp.yw - (double)wp.ry);
                float gx = 0.f, gy = 0.f;
                if (wp.rx >= 1 && wp.rx <= W - 2 && wp.ry >= 1 && wp.ry <= H - 2) {
                    const float* Dp = dldi32 + (int64_t)r * HW + (int64_t)wp.ry * W + wp.rx;
                    float d[9];
#pragma unroll
                    for (int j = -1; j <= 1; ++j)
#pragma unroll
                        for (int i = -1; i <= 1; ++i) d[(j + 1) * 3 + (i + 1)] = __ldg(Dp + j * W + i);
                    const float A = __expf(-0.5f * (fx * fx + fy * fy));
                    const float wx0 = kG1 * __expf(-fx), wx2 = kG1 * __expf(fx);          // wx1 = 1
                    const float wy0 = kG1 * __expf(-fy), wy2 = kG1 * __expf(fy);          // wy1 = 1
                    const float ux0 = wx0 * (-1.f - fx), ux1 = -fx, ux2 = wx2 * (1.f - fx);  // wx_i (i - fx)
                    float sj[3], sxj[3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        sj[j] = fmaf(d[j * 3 + 0], wx0, fmaf(d[j * 3 + 2], wx2, d[j * 3 + 1]));
                        sxj[j] = fmaf(d[j * 3 + 0], ux0, fmaf(d[j * 3 + 2], ux2, d[j * 3 + 1] * ux1));
                    }
                    gx = A * fmaf(wy0, sxj[0], fmaf(wy2, sxj[2], sxj[1]));
                    gy = A * fmaf(wy0 * (-1.f - fy), sj[0], fmaf(wy2 * (1.f - fy), sj[2], -fy * sj[1]));
                } else {
                    const double* img = dldi + (int64_t)r * HW;
#pragma unroll
                    for (int i = -1; i <= 1; ++i) {
#pragma unroll
                        for (int j = -1; j <= 1; ++j) {
                            int rr = wp.ry + j, cc = wp.rx + i;
                            if (drop_index<WRAP>(rr, cc, H, W)) {
                                const float qx = (float)i - fx, qy = (float)j - fy;
                                const float g = (float)img[rr * W + cc] * (__expf(-0.5f * (qx * qx + qy * qy)) * (float)kInv2Pi);
                                gx = fmaf(g, qx, gx);
                                gy = fmaf(g, qy, gy);
                            }
                        }
                    }
                }
                const float dtf = (float)dt;
                gx_acc = fmaf(-dtf, gx, gx_acc);
                gy_acc = fmaf(-dtf, gy, gy_acc);
            }
        }
        // segmented (by source pixel) inclusive suffix sum inside the warp (float32: <= 32 terms), one float64 RED per run
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float ox = __shfl_down_sync(0xffffffffu, gx_acc, o);
            const float oy = __shfl_down_sync(0xffffffffu, gy_acc, o);
            const uint32_t oxy = __shfl_down_sync(0xffffffffu, xy, o);
            if (lane + o < 32 && oxy == xy) { gx_acc += ox; gy_acc += oy; }
        }
        const uint32_t prev = __shfl_up_sync(0xffffffffu, xy, 1);
        const bool head = (lane == 0) || (prev != xy);
        if (head && e < n) {
            const int x = xy & 0xffffu, y = xy >> 16;
            atomicAdd(&G[(y * W + x) * 2 + 0], (double)gx_acc);
            atomicAdd(&G[(y * W + x) * 2 + 1], (double)gy_acc);
        }
    }
}

}  // namespace eincm
