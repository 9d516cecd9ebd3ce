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

// exp(x) for |x| < ~1 (no range guards needed): one FMUL + MUFU.EX2
__device__ __forceinline__ float exp_small(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}

// moments m[(j+1)*3 + (i+1)] = A * Bx^i * By^j for the sub-pixel offset (fx, fy)
__device__ __forceinline__ void moments9(float fx, float fy, float m[9]) {
    const float A = exp_small(-0.5f * (fx * fx + fy * fy));
    const float bx = exp_small(fx), ibx = exp_small(-fx), by = exp_small(fy), iby = exp_small(-fy);
    const float r0 = A * iby, r2 = A * by;
    m[0] = r0 * ibx; m[1] = r0; m[2] = r0 * bx;
    m[3] = A * ibx;  m[4] = A;  m[5] = A * bx;
    m[6] = r2 * ibx; m[7] = r2; m[8] = r2 * bx;
}

// Warp of one event to one reference time on the fast path.  Same float64 arithmetic, in the same order, as warp_event
// (event_warpers.py:34-35: x' = x - (theta * dt) * 1.0), but rint() and the int conversion use the 2^52 magic constant
// (two DADDs instead of F2I + I2F on the slow conversion pipe): for |x'| < 2^31, (x' + M) - M == rint(x') under
// round-half-to-even and the low word of (x' + M) is that integer.  `cls`: 0 = dropped (non-finite / absurdly far),
// 1 = centre is an interior pixel (fast record path), 2 = border / outside (per-tap index rule).
struct Hit { int rx, ry; float fx, fy; int cls; };

__device__ __forceinline__ Hit warp_hit(double xd, double yd, double thx, double thy, double dt, int H, int W) {
    constexpr double kMagic = 6755399441055744.0;   // 1.5 * 2^52
    const double xw = __dsub_rn(xd, __dmul_rn(thx, dt));
    const double yw = __dsub_rn(yd, __dmul_rn(thy, dt));
    const bool ok = (fabs(xw) < 1.0e9) && (fabs(yw) < 1.0e9);      // false for NaN
    const double sx = __dadd_rn(xw, kMagic), sy = __dadd_rn(yw, kMagic);
    Hit h;
    h.rx = __double2loint(sx);
    h.ry = __double2loint(sy);
    h.fx = (float)__dsub_rn(xw, __dsub_rn(sx, kMagic));
    h.fy = (float)__dsub_rn(yw, __dsub_rn(sy, kMagic));
    const bool interior = ((unsigned)(h.rx - 1) < (unsigned)(W - 2)) && ((unsigned)(h.ry - 1) < (unsigned)(H - 2));
    h.cls = ok ? (interior ? 1 : 2) : 0;
    return h;
}

// ---- forward: warp + moment splat ------------------------------------------------------------------------------
// Event stream layout (eincm_plan_set_window): sorted by source pixel (tile-major key) and, inside a pixel, by time; padded
// with kNoEvent sentinels to a multiple of kEvK.  One thread owns kEvK CONSECUTIVE events (one 128-bit load of packed
// coordinates, two of timestamps) and warps each of them to RB reference times at once (RB independent dependency chains).
// Consecutive events are mostly the same pixel a little later, so their warped centres often coincide: per reference time
// the thread keeps a running record (9 moments + destination key) in registers and only issues the three reductions when the
// destination changes (run-length pre-aggregation; worst case = one flush per event and reference).
// theta_full == nullptr: zero flow (the un-warped image of events, losses.py:54).
constexpr int kEvK = 4;
constexpr uint32_t kNoEvent = 0xffffffffu;
constexpr int kMaxRB = 4;      // reference times processed per pass (register budget: RB * 10 accumulators)

__device__ __forceinline__ void flush9(float* __restrict__ Cr, int key, const float a[9]) {
    float* rec = Cr + (int64_t)key * kRec;
    red_add_v4(rec, a[0], a[1], a[2], a[3]);
    red_add_v4(rec + 4, a[4], a[5], a[6], a[7]);
    red_add_f32(rec + 8, a[8]);
}

// slow path of one event whose centre is not an interior pixel: the reference's per-tap index rule; the value 2 pi v_ij goes
// to channel (0,0) (index 4) of the destination record
template <bool WRAP>
__device__ __noinline__ void splat9_border(int rx, int ry, float fx, float fy, int H, int W, float* __restrict__ Cr) {
#pragma unroll
    for (int i = -1; i <= 1; ++i) {
#pragma unroll
        for (int j = -1; j <= 1; ++j) {
            int rr = ry + j, cc = rx + i;
            if (drop_index<WRAP>(rr, cc, H, W)) {
                const float qx = (float)i - fx, qy = (float)j - fy;
                red_add_f32(Cr + ((int64_t)rr * W + cc) * kRec + 4, __expf(-0.5f * (qx * qx + qy * qy)));
            }
        }
    }
}

struct EventGroup {
    uint32_t xy[kEvK];
    double t[kEvK];
};

__device__ __forceinline__ void load_group(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, int64_t g, EventGroup& G) {
    static_assert(kEvK == 4, "one uint4 of packed coordinates per thread");
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(ev_xy) + g);
    const double2 ta = __ldg(reinterpret_cast<const double2*>(ev_t) + 2 * g);
    const double2 tb = __ldg(reinterpret_cast<const double2*>(ev_t) + 2 * g + 1);
    G.xy[0] = q.x; G.xy[1] = q.y; G.xy[2] = q.z; G.xy[3] = q.w;
    G.t[0] = ta.x; G.t[1] = ta.y; G.t[2] = tb.x; G.t[3] = tb.y;
}

template <bool WRAP, int RB>
__global__ void __launch_bounds__(256)
k_splat9(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, int64_t n_groups,
         const double2* __restrict__ theta_full, int H, int W, int R, const __grid_constant__ RefTimes tref,
         float* __restrict__ C /* [R][H*W][kRec] */) {
    const int64_t HW = (int64_t)H * W;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_groups; g += (int64_t)gridDim.x * blockDim.x) {
        EventGroup ev;
        load_group(ev_xy, ev_t, g, ev);
        double2 th[kEvK];
#pragma unroll
        for (int k = 0; k < kEvK; ++k) {
            th[k] = make_double2(0.0, 0.0);
            if (theta_full != nullptr && ev.xy[k] != kNoEvent)
                th[k] = __ldg(theta_full + (int)(ev.xy[k] >> 16) * W + (int)(ev.xy[k] & 0xffffu));
        }
        for (int r0 = 0; r0 < R; r0 += RB) {
            float acc[RB][9];
            int acc_key[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) acc_key[r] = -1;
#pragma unroll
            for (int k = 0; k < kEvK; ++k) {
                if (ev.xy[k] == kNoEvent) continue;       // only in the last group of the stream
                const double xd = (double)(ev.xy[k] & 0xffffu), yd = (double)(ev.xy[k] >> 16);
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    if (r0 + r >= R) continue;            // uniform
                    float* Cr = C + (int64_t)(r0 + r) * HW * kRec;
                    const Hit h = warp_hit(xd, yd, th[k].x, th[k].y, ev.t[k] - tref.t[r0 + r], H, W);
                    if (h.cls == 1) {
                        float m[9];
                        moments9(h.fx, h.fy, m);
                        const int key = h.ry * W + h.rx;
                        if (key == acc_key[r]) {
#pragma unroll
                            for (int c = 0; c < 9; ++c) acc[r][c] += m[c];
                        } else {
                            if (acc_key[r] >= 0) flush9(Cr, acc_key[r], acc[r]);
#pragma unroll
                            for (int c = 0; c < 9; ++c) acc[r][c] = m[c];
                            acc_key[r] = key;
                        }
                    } else if (h.cls == 2) {
                        splat9_border<WRAP>(h.rx, h.ry, h.fx, h.fy, H, W, Cr);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RB; ++r)
                if (r0 + r < R && acc_key[r] >= 0) flush9(C + (int64_t)(r0 + r) * HW * kRec, acc_key[r], acc[r]);
        }
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
// Same thread-owns-kEvK-consecutive-events layout as k_splat9, RB reference times per event in flight; per-event results are
// first merged per source pixel inside the thread, then across the warp.
template <bool WRAP>
__device__ __noinline__ float2 backward9_border(int rx, int ry, float fx, float fy, int H, int W, const double* __restrict__ img) {
    float gx = 0.f, gy = 0.f;
#pragma unroll
    for (int i = -1; i <= 1; ++i) {
#pragma unroll
        for (int j = -1; j <= 1; ++j) {
            int rr = ry + j, cc = rx + i;
            if (drop_index<WRAP>(rr, cc, H, W)) {
                const float qx = (float)i - fx, qy = (float)j - fy;
                const float g = (float)img[rr * W + cc] * (__expf(-0.5f * (qx * qx + qy * qy)) * (float)kInv2Pi);
                gx = fmaf(g, qx, gx);
                gy = fmaf(g, qy, gy);
            }
        }
    }
    return make_float2(gx, gy);
}

__device__ __forceinline__ void red_G(double* __restrict__ G, int W, uint32_t xy, float sx, float sy) {
    double* g = G + ((int64_t)(xy >> 16) * W + (xy & 0xffffu)) * 2;
    atomicAdd(g, (double)sx);
    atomicAdd(g + 1, (double)sy);
}

template <bool WRAP, int RB>
__global__ void __launch_bounds__(256)
k_backward9(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, int64_t n_groups,
            const double2* __restrict__ theta_full, int H, int W, int R, const __grid_constant__ RefTimes tref,
            const float* __restrict__ dldi32 /* [R][H][W], already scaled by 1/(2 pi) */, const double* __restrict__ dldi,
            double* __restrict__ G /* [H][W][2] */) {
    const int64_t HW = (int64_t)H * W;
    const int lane = threadIdx.x & 31;
    const int64_t n_round = (n_groups + 31) / 32 * 32;     // whole warps stay in the loop for the shuffles
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_round; g += (int64_t)gridDim.x * blockDim.x) {
        float ax[kEvK], ay[kEvK];
        EventGroup ev;
#pragma unroll
        for (int k = 0; k < kEvK; ++k) { ax[k] = 0.f; ay[k] = 0.f; ev.xy[k] = kNoEvent; }
        if (g < n_groups) {
            load_group(ev_xy, ev_t, g, ev);
#pragma unroll
            for (int k = 0; k < kEvK; ++k) {
                if (ev.xy[k] == kNoEvent) continue;
                const int x = ev.xy[k] & 0xffffu, y = ev.xy[k] >> 16;
                const double2 th = __ldg(theta_full + y * W + x);
                const double xd = (double)x, yd = (double)y;
                const float tf = (float)ev.t[k];
                for (int r0 = 0; r0 < R; r0 += RB) {
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        if (r0 + r >= R) continue;        // uniform
                        const double tr = tref.t[r0 + r];
                        const Hit h = warp_hit(xd, yd, th.x, th.y, ev.t[k] - tr, H, W);
                        float gx = 0.f, gy = 0.f;
                        if (h.cls == 1) {
                            const float* Dp = dldi32 + (int64_t)(r0 + r) * HW + (h.ry * W + h.rx);
                            float d[9];
#pragma unroll
                            for (int j = -1; j <= 1; ++j)
#pragma unroll
                                for (int i = -1; i <= 1; ++i) d[(j + 1) * 3 + (i + 1)] = __ldg(Dp + j * W + i);
                            const float fx = h.fx, fy = h.fy;
                            const float A = exp_small(-0.5f * (fx * fx + fy * fy));
                            const float wx0 = kG1 * exp_small(-fx), wx2 = kG1 * exp_small(fx);          // wx1 = 1
                            const float wy0 = kG1 * exp_small(-fy), wy2 = kG1 * exp_small(fy);          // wy1 = 1
                            const float ux0 = wx0 * (-1.f - fx), ux1 = -fx, ux2 = wx2 * (1.f - fx);     // wx_i (i - fx)
                            float sj[3], sxj[3];
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                sj[j] = fmaf(d[j * 3 + 0], wx0, fmaf(d[j * 3 + 2], wx2, d[j * 3 + 1]));
                                sxj[j] = fmaf(d[j * 3 + 0], ux0, fmaf(d[j * 3 + 2], ux2, d[j * 3 + 1] * ux1));
                            }
                            gx = A * fmaf(wy0, sxj[0], fmaf(wy2, sxj[2], sxj[1]));
                            gy = A * fmaf(wy0 * (-1.f - fy), sj[0], fmaf(wy2 * (1.f - fy), sj[2], -fy * sj[1]));
                        } else if (h.cls == 2) {
                            const float2 gb = backward9_border<WRAP>(h.rx, h.ry, h.fx, h.fy, H, W, dldi + (int64_t)(r0 + r) * HW);
                            gx = gb.x; gy = gb.y;
                        }
                        const float dtf = tf - (float)tr;      // float32 weight of a float32 result (|dt| ~ 1)
                        ax[k] = fmaf(-dtf, gx, ax[k]);
                        ay[k] = fmaf(-dtf, gy, ay[k]);
                    }
                }
            }
        }
        // per-thread runs of equal source pixel: all but the last run go straight to G
        uint32_t run_xy = ev.xy[0];
        float sx = ax[0], sy = ay[0];
#pragma unroll
        for (int k = 1; k < kEvK; ++k) {
            if (ev.xy[k] == run_xy) { sx += ax[k]; sy += ay[k]; }
            else {
                if (run_xy != kNoEvent) red_G(G, W, run_xy, sx, sy);
                run_xy = ev.xy[k]; sx = ax[k]; sy = ay[k];
            }
        }
        // the last runs of the warp's threads: segmented (by source pixel) inclusive suffix sum, one RED pair per run.
        // Keys are sorted, so lane+o holding the same last key means every event in between has that key.
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float ox = __shfl_down_sync(0xffffffffu, sx, o);
            const float oy = __shfl_down_sync(0xffffffffu, sy, o);
            const uint32_t oxy = __shfl_down_sync(0xffffffffu, run_xy, o);
            if (lane + o < 32 && oxy == run_xy) { sx += ox; sy += oy; }
        }
        const uint32_t prev = __shfl_up_sync(0xffffffffu, run_xy, 1);
        const bool head = (lane == 0) || (prev != run_xy);
        if (head && run_xy != kNoEvent) red_G(G, W, run_xy, sx, sy);
    }
}

// ---- debug tap: Xs_rounded (event_utils.py:33) of reference r, written back in the original event order ----
// EXACT selects the conversion used by the float64 nine-tap kernels (cvt.rni), otherwise the magic-constant rint of the moment
// kernels: the tap reports the indices the active kernels really use.
template <bool EXACT>
__global__ void k_rounded_pixels(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, const uint32_t* __restrict__ perm,
                                 int64_t n, const double2* __restrict__ theta_full, int H, int W, double t_ref,
                                 int32_t* __restrict__ cols_out, int32_t* __restrict__ rows_out) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t xy = ev_xy[e];
        if (xy == kNoEvent) continue;
        const int x = xy & 0xffffu, y = xy >> 16;
        const double2 th = theta_full[y * W + x];
        const uint32_t o = perm[e];
        if (EXACT) {
            const Warped wp = warp_event(x, y, th.x, th.y, ev_t[e] - t_ref);
            cols_out[o] = wp.ok ? wp.rx : INT32_MAX;
            rows_out[o] = wp.ok ? wp.ry : INT32_MAX;
        } else {
            const Hit h = warp_hit((double)x, (double)y, th.x, th.y, ev_t[e] - t_ref, H, W);
            cols_out[o] = h.cls ? h.rx : INT32_MAX;
            rows_out[o] = h.cls ? h.ry : INT32_MAX;
        }
    }
}

}  // namespace eincm
