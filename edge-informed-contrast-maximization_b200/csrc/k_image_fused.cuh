// Fused image-space pass of the default (float32 moment splat, delta == 0, single GPU) path: two kernels per evaluation
// instead of compose + A + scalars + B + scalars + C.
//
//   k_img_fused1: moment records -> image of warped events (same arithmetic as k_compose9), Scharr contrast, min / max with
//                 tie counts, and the moments  sum I, sum I^2, sum E*I  from which the min-max-normalised MSE and the sums of
//                 its backward follow algebraically once the GLOBAL min / max are known:
//                     N = (I - m)/D,  D = max - m + eps
//                     sum (E-N)^2     = sum E^2 - 2/D (sum EI - m sum E) + Q/D^2,      Q = sum I^2 - 2 m sum I + HW m^2
//                     s1 = sum gN     = cb (sum E - (sum I - HW m)/D)                   gN = cb (E - N)
//                     s2 = sum gN(I-m)= cb ((sum EI - m sum E) - Q/D)
//                 so no second pass over the images is needed (reference: src/utils/img_utils.py:24-25,
//                 src/eincm/objectives/correlation_objectives.py:25-26, contrast_objectives.py:22-25).  The last CTA to
//                 finish reduces the per-CTA partials in a fixed order and evaluates the loss (src/eincm/losses.py:171-193).
//   k_img_fused3: d loss / d IWE (float64 + float32/2pi copies), as k_img_C.
#pragma once
#include "common.cuh"
#include "k_events9.cuh"
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

// block-wide merge of FusedAcc (kFNT threads); result valid in thread 0.  Fixed order: deterministic.
__device__ __forceinline__ FusedAcc fused_block_reduce(FusedAcc a, double (*sh)[kFPart]) {
    const int tid = linear_tid();
#pragma unroll
    for (int o = 16; o; o >>= 1) a.merge(a.shfl_xor(o));
    __syncthreads();     // sh may still be read from a previous call
    if ((tid & 31) == 0) {
        double* d = sh[tid >> 5];
        d[0] = a.sq; d[1] = a.sI; d[2] = a.sI2; d[3] = a.sEI; d[4] = a.mn; d[5] = a.cmn; d[6] = a.mx; d[7] = a.cmx;
    }
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < kFNT / 32; ++k) {
            FusedAcc o;
            o.sq = sh[k][0]; o.sI = sh[k][1]; o.sI2 = sh[k][2]; o.sEI = sh[k][3]; o.mn = sh[k][4]; o.cmn = sh[k][5]; o.mx = sh[k][6]; o.cmx = sh[k][7];
            a.merge(o);
        }
    }
    return a;
}

// grid (tiles_x, tiles_y, R), block (kFTX, 8): each thread owns 2 pixels of a 32x16 tile.
// FIX: the source is the fixed-point image of the tile-privatised splat (k_events_tile.cuh), converted exactly to float64;
// otherwise the float32 moment records of k_splat9, composed as in k_compose9.
template <bool FIX>
__global__ void __launch_bounds__(kFNT)
k_img_fused1(const void* __restrict__ src, const double* __restrict__ edges, int H, int W, int nb /* tiles per image */,
             double* __restrict__ iwe, double* __restrict__ part /* [R][nb][kFPart] */, DevScalars* sc,
             double alpha, double beta, double gamma, int use_tv, double* __restrict__ loss_out) {
    constexpr int RW = kFTX + 4, RH = kFTY + 4;     // record cells (halo 2)
    constexpr int IW = kFTX + 2, IH = kFTY + 2;     // image cells (halo 1)
    __shared__ float rec[FIX ? 1 : RH][FIX ? 1 : RW][9];
    __shared__ double img[IH][IW];
    __shared__ double red[kFNT / 32][kFPart];
    __shared__ bool last;
    const int r = blockIdx.z, R = gridDim.z;
    const int64_t HW = (int64_t)H * W;
    const int x0 = blockIdx.x * kFTX, y0 = blockIdx.y * kFTY;
    const int tid = linear_tid();
    if (FIX) {
        const unsigned long long* Fr = reinterpret_cast<const unsigned long long*>(src) + (int64_t)r * HW;
        constexpr int NIT = (IW * IH + kFNT - 1) / kFNT;
        unsigned long long v[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int k = tid + it * kFNT;
            const int cy = k / IW, cx = k % IW;
            const int X = x0 + cx - 1, Y = y0 + cy - 1;
            v[it] = (k < IW * IH && X >= 0 && X < W && Y >= 0 && Y < H) ? __ldcg(Fr + (int64_t)Y * W + X) : 0ull;
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int k = tid + it * kFNT;
            if (k < IW * IH) (&img[0][0])[k] = (double)(long long)v[it] * kFixToIwe;
        }
    } else {
    const float* Cr = reinterpret_cast<const float*>(src) + (int64_t)r * HW * kRec;
    {
        // all global loads first (independent, in flight together), then the shared-memory stores
        constexpr int NIT = (RW * RH * 3 + kFNT - 1) / kFNT;
        float4 v[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int k = tid + it * kFNT;
            const int p3 = k % 3, cell = k / 3;
            const int ly = cell / RW, lx = cell % RW;
            const int yy = y0 + ly - 2, xx = x0 + lx - 2;
            v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < RW * RH * 3 && xx >= 0 && xx < W && yy >= 0 && yy < H)
                v[it] = __ldcg(reinterpret_cast<const float4*>(Cr + ((int64_t)yy * W + xx) * kRec) + p3);
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int k = tid + it * kFNT;
            if (k < RW * RH * 3) {
                const int p3 = k % 3, cell = k / 3;
                float* dst = &rec[0][0][0] + cell * 9 + 4 * p3;
                dst[0] = v[it].x;
                if (p3 < 2) { dst[1] = v[it].y; dst[2] = v[it].z; dst[3] = v[it].w; }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = tid; k < IW * IH; k += kFNT) {
        const int cy = k / IW, cx = k % IW;
        const int X = x0 + cx - 1, Y = y0 + cy - 1;
        double v = 0.0;
        if (X >= 0 && X < W && Y >= 0 && Y < H) {
            const int ly = cy + 1, lx = cx + 1;      // this cell in record coordinates
            double corners = 0.0, edg = 0.0;         // same order as k_compose9
            corners += (double)rec[ly + 1][lx + 1][0];
            corners += (double)rec[ly + 1][lx - 1][2];
            corners += (double)rec[ly - 1][lx + 1][6];
            corners += (double)rec[ly - 1][lx - 1][8];
            edg += (double)rec[ly + 1][lx][1];
            edg += (double)rec[ly][lx + 1][3];
            edg += (double)rec[ly][lx - 1][5];
            edg += (double)rec[ly - 1][lx][7];
            const double centre = (double)rec[ly][lx][4];
            constexpr double g1 = 0.60653065971263342, g2 = 0.36787944117144233;
            v = (centre + g1 * edg + g2 * corners) * kInv2Pi;
        }
        img[cy][cx] = v;
    }
    }
    __syncthreads();
    FusedAcc acc;
    acc.init();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int ty = threadIdx.y + 8 * h;
        const int x = x0 + threadIdx.x, y = y0 + ty;
        if (x < W && y < H) {
            const double* p = &img[ty + 1][threadIdx.x + 1];
            double gx, gy;
            scharr_at(p, IW, gx, gy);
            const double I = *p;
            const int64_t q = (int64_t)y * W + x;
            const double E = edges[(int64_t)r * HW + q];
            iwe[(int64_t)r * HW + q] = I;
            FusedAcc o;
            o.sq = gx * gx + gy * gy; o.sI = I; o.sI2 = I * I; o.sEI = E * I; o.mn = I; o.cmn = 1.0; o.mx = I; o.cmx = 1.0;
            acc.merge(o);
        }
    }
    acc = fused_block_reduce(acc, red);
    const int b = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) {
        double* d = part + ((int64_t)r * nb + b) * kFPart;
        d[0] = acc.sq; d[1] = acc.sI; d[2] = acc.sI2; d[3] = acc.sEI; d[4] = acc.mn; d[5] = acc.cmn; d[6] = acc.mx; d[7] = acc.cmx;
    }
    if (!last_block_ticket(&sc->counters[4], (unsigned)(nb * R), &last)) return;
    // ---- last CTA: global statistics per reference image, cotangent scales, loss -------------------------------------
    if (tid == 0) scalars_coefs(sc, R, (double)HW, alpha, beta, 0.0, 0);
    __syncthreads();
    for (int q = 0; q < R; ++q) {
        FusedAcc a;
        a.init();
        for (int k = tid; k < nb; k += kFNT) {
            const double* d = part + ((int64_t)q * nb + k) * kFPart;
            FusedAcc o;
            o.sq = __ldcg(d + 0); o.sI = __ldcg(d + 1); o.sI2 = __ldcg(d + 2); o.sEI = __ldcg(d + 3);
            o.mn = __ldcg(d + 4); o.cmn = __ldcg(d + 5); o.mx = __ldcg(d + 6); o.cmx = __ldcg(d + 7);
            a.merge(o);
        }
        a = fused_block_reduce(a, red);
        if (tid == 0) {
            const double n = (double)HW;
            const double m = a.mn, D = (a.mx - a.mn) + kEps;                    // img_utils.py:25
            const double sE = sc->sumE[q], sE2 = sc->sumE2[q];
            const double EIm = a.sEI - m * sE;                                   // sum E (I - m)
            const double Q = a.sI2 - 2.0 * m * a.sI + n * m * m;                 // sum (I - m)^2
            const double cb = sc->coefB[q];
            Stats& st = sc->ref[q];
            st.contrast = a.sq / n;
            st.mn = m; st.mx = a.mx; st.D = D;
            st.mse = (sE2 - 2.0 * EIm / D + Q / (D * D)) / n;
            st.s1 = cb * (sE - (a.sI - n * m) / D);
            st.s2 = cb * (EIm - Q / D);
            st.cnt_min = a.cmn; st.cnt_max = a.cmx;
            st.div = 0.0;
        }
    }
    __syncthreads();
    if (tid == 0) {
        scalars_loss(sc, R, alpha, beta, gamma, 0.0, use_tv, 0, loss_out);
        sc->counters[4] = 0;
    }
}

// per-window: sum E_r and sum E_r^2 (deterministic single-CTA-per-reference reduction; once per window)
__global__ void __launch_bounds__(1024)
k_edge_sums(const double* __restrict__ edges, int64_t HW, DevScalars* sc) {
    __shared__ double sh[32];
    const int r = blockIdx.x;
    double s = 0.0, s2 = 0.0;
    for (int64_t p = threadIdx.x; p < HW; p += blockDim.x) { const double e = edges[r * HW + p]; s += e; s2 += e * e; }
    s = block_reduce<1024>(s, OpSum(), sh);
    s2 = block_reduce<1024>(s2, OpSum(), sh);
    if (threadIdx.x == 0) { sc->sumE[r] = s; sc->sumE2[r] = s2; }
}

// dLdI_r = coefA[r] * (corr2d(Gx,Kx) + corr2d(Gy,Ky)) + gN/D + g_m [I==min]/#min + g_M [I==max]/#max   (as k_img_C)
__global__ void __launch_bounds__(kFNT)
k_img_fused3(const double* __restrict__ imgs, const double* __restrict__ edges, int H, int W, const DevScalars* __restrict__ sc,
             double* __restrict__ dldi, float* __restrict__ dldi32) {
    constexpr int PW = kFTX + 4, PH = kFTY + 4, GW = kFTX + 2, GH = kFTY + 2;
    __shared__ double tile[PH][PW];
    __shared__ double gxs[GH][GW], gys[GH][GW];
    const int r = blockIdx.z;
    const int64_t HW = (int64_t)H * W;
    const double* img = imgs + r * HW;
    const int x0 = blockIdx.x * kFTX, y0 = blockIdx.y * kFTY;
    const int tid = linear_tid();
    {
        constexpr int NIT = (PW * PH + kFNT - 1) / kFNT;
        double v[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int k = tid + it * kFNT;
            const int ly = k / PW, lx = k % PW;
            const int y = y0 + ly - 2, x = x0 + lx - 2;
            v[it] = (k < PW * PH && x >= 0 && x < W && y >= 0 && y < H) ? img[(int64_t)y * W + x] : 0.0;
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int k = tid + it * kFNT;
            if (k < PW * PH) (&tile[0][0])[k] = v[it];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = tid; k < GW * GH; k += kFNT) {
        const int ly = k / GW, lx = k % GW;
        const int y = y0 + ly - 1, x = x0 + lx - 1;
        double gx = 0.0, gy = 0.0;
        if (x >= 0 && x < W && y >= 0 && y < H) scharr_at(&tile[ly + 1][lx + 1], PW, gx, gy);
        gxs[ly][lx] = gx; gys[ly][lx] = gy;
    }
    __syncthreads();
    const Stats st = sc->ref[r];
    const double cA = sc->coefA[r], cB = sc->coefB[r];
    const double g_M = -st.s2 / (st.D * st.D);
    const double g_m = -st.s1 / st.D + st.s2 / (st.D * st.D);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int ty = threadIdx.y + 8 * h;
        const int x = x0 + threadIdx.x, y = y0 + ty;
        if (x >= W || y >= H) continue;
        const double I = tile[ty + 2][threadIdx.x + 2];
        const int64_t p = (int64_t)y * W + x;
        const double adj = scharr_adjoint_at(&gxs[ty + 1][threadIdx.x + 1], &gys[ty + 1][threadIdx.x + 1], GW);
        const double c = I - st.mn;
        const double gN = cB * (edges[r * HW + p] - c / st.D);
        double out = cA * adj + gN / st.D;
        if (I == st.mn) out += g_m / st.cnt_min;
        if (I == st.mx) out += g_M / st.cnt_max;
        dldi[r * HW + p] = out;
        dldi32[r * HW + p] = (float)(out * kInv2Pi);
    }
}

}  // namespace eincm
