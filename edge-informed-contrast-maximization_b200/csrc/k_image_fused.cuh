// Fused image-space pass of the default (tile-privatised splat, delta == 0, single GPU) path: two kernels per evaluation
// instead of compose + A + scalars + B + scalars + C.
//
//   k_image_pass: fixed-point image -> float64 image of warped events, Scharr contrast, min / max with
//                 tie counts, and the moments  sum I, sum I^2, sum E*I  from which the min-max-normalised MSE and the sums of
//                 its backward follow algebraically once the GLOBAL min / max are known:
//                     N = (I - m)/D,  D = max - m + eps
//                     sum (E-N)^2     = sum E^2 - 2/D (sum EI - m sum E) + Q/D^2,      Q = sum I^2 - 2 m sum I + HW m^2
//                     s1 = sum gN     = cb (sum E - (sum I - HW m)/D)                   gN = cb (E - N)
//                     s2 = sum gN(I-m)= cb ((sum EI - m sum E) - Q/D)
//                 so no second pass over the images is needed for the statistics (reference: src/utils/img_utils.py:24-25,
//                 src/eincm/objectives/correlation_objectives.py:25-26, contrast_objectives.py:22-25, src/eincm/losses.py:171-193);
//                 then d loss / d IWE (float64 + float32/2pi copies), as k_img_C.
#pragma once
#include <cooperative_groups.h>

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

// ---- the cooperative image pass -------------------------------------------------------------------------------------------
// One persistent cooperative kernel (grid = min(2 x SMs, tiles), all CTAs co-resident) does the whole image-space part of an
// evaluation:
//   phase 1  every CTA walks its contiguous share of the R x tiles work list: fixed-point image -> float64 image (written out
//            for phase 3 and for the debug taps), the fixed-point cells are cleared for the next evaluation (no memset launch),
//            per-thread statistics are accumulated over all its pixels and reduced ONCE per (CTA, reference image);
//   barrier  grid-wide;
//   phase 2  every CTA reduces the per-CTA partials of the reference images it owns in the same fixed order (identical,
//            deterministic results everywhere - no serial "last CTA" tail); CTA 0 also evaluates the loss;
//   phase 3  d loss / d IWE of the CTA's tiles.
constexpr int kCoopMaxRefs = EINCM_MAX_REFS;

struct ImagePassArgs {
    unsigned long long* fix;      // [R][H*W] fixed-point images of warped events (read, then cleared)
    const double* edges;          // [R][H*W]
    double* iwe;                  // [R][H*W] float64 images (written in phase 1, read in phase 3)
    double* part;                 // [R][grid][kFPart]
    DevScalars* sc;
    double* dldi;                 // [R][H*W] float64 d loss / d IWE (debug tap) or null
    float* dldi32;                // [R][H*W] float32 d loss / d IWE / (2 pi)
    double* loss_out;
    double* zero_buf;             // buffers cleared for the event backward pass (dense flow-field gradient, theta gradient) or null
    double* zero_buf2;
    int n_zero, n_zero2;
    int H, W, R, tiles_x, tiles_y;
    double alpha, beta, gamma;
    int use_tv, want_grad;
};

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

// 8-byte asynchronous global -> shared copy; `valid == false` zero-fills (image border)
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kStages = 3;                          // tiles in flight per CTA (cp.async ring)

struct ImagePassSmem {
    // phase 1: raw fixed-point tile (halo 1), converted in place to float64; phase 3: float64 tile (halo 2)
    double img[kStages][(kFTY + 4) * (kFTX + 4)];
    double edg[kStages][kFTY * kFTX];
    double gxs[(kFTY + 2) * (kFTX + 2)], gys[(kFTY + 2) * (kFTX + 2)];
    double red[kFNT / 32][kFPart];
    Stats st[kCoopMaxRefs];
    double coefA[kCoopMaxRefs], coefB[kCoopMaxRefs];
};

// issue the asynchronous loads of one tile: HALO cells around the kFTX x kFTY interior of image `src` (8-byte cells), and
// the tile's edge-image cells
template <int HALO>
__device__ __forceinline__ void issue_tile(const void* __restrict__ src, const double* __restrict__ edges, int H, int W, int x0, int y0,
                                           double* __restrict__ img_dst, double* __restrict__ edg_dst) {
    constexpr int TW = kFTX + 2 * HALO, TH = kFTY + 2 * HALO;
    const int tid = linear_tid();
    const unsigned long long* s8 = reinterpret_cast<const unsigned long long*>(src);
    for (int k = tid; k < TW * TH; k += kFNT) {
        const int cy = k / TW, cx = k - cy * TW;
        const int X = x0 + cx - HALO, Y = y0 + cy - HALO;
        const bool ok = (X >= 0) & (X < W) & (Y >= 0) & (Y < H);
        cp_async8(img_dst + k, ok ? (const void*)(s8 + (int64_t)Y * W + X) : (const void*)s8, ok);
    }
    for (int k = tid; k < kFTX * kFTY; k += kFNT) {
        const int cy = k / kFTX, cx = k - cy * kFTX;
        const int X = x0 + cx, Y = y0 + cy;
        const bool ok = (X < W) & (Y < H);
        cp_async8(edg_dst + k, ok ? (const void*)(edges + (int64_t)Y * W + X) : (const void*)edges, ok);
    }
}

__global__ void __launch_bounds__(kFNT)
k_image_pass(const ImagePassArgs A) {
    constexpr int IW = kFTX + 2;                    // row pitch of a halo-1 tile (phase 1)
    constexpr int PW = kFTX + 4;                    // row pitch of a halo-2 tile (phase 3)
    constexpr int GW = kFTX + 2, GH = kFTY + 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ImagePassSmem& S = *reinterpret_cast<ImagePassSmem*>(smem_raw);
    const int H = A.H, W = A.W, R = A.R;
    const int64_t HW = (int64_t)H * W;
    const int tid = linear_tid();
    const int tpi = A.tiles_x * A.tiles_y;          // tiles per image
    const int T = tpi * R;
    const int G = gridDim.x, b = blockIdx.x;
    const int t_begin = (int)(((long long)T * b) / G), t_end = (int)(((long long)T * (b + 1)) / G);
    const int r_first = t_begin < t_end ? t_begin / tpi : 0, r_last = t_begin < t_end ? (t_end - 1) / tpi : -1;
    auto tile_origin = [&](int t, int& r, int& x0, int& y0) {
        r = t / tpi;
        const int tt = t - r * tpi;
        const int ty = tt / A.tiles_x;
        x0 = (tt - ty * A.tiles_x) * kFTX; y0 = ty * kFTY;
    };

    // identity partials for every reference image (a CTA usually touches one or two)
    for (int k = tid; k < R * kFPart; k += kFNT) {
        const int q = k / kFPart, f = k % kFPart;
        A.part[((int64_t)q * G + b) * kFPart + f] = (f == 4) ? INFINITY : ((f == 6) ? -INFINITY : 0.0);
    }
    if (A.zero_buf != nullptr)
        for (int k = b * kFNT + tid; k < A.n_zero; k += G * kFNT) A.zero_buf[k] = 0.0;
    if (A.zero_buf2 != nullptr)
        for (int k = b * kFNT + tid; k < A.n_zero2; k += G * kFNT) A.zero_buf2[k] = 0.0;

    // ---- phase 1: statistics ----------------------------------------------------------------------------------------------
    FusedAcc acc;
    acc.init();
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
        if (t_begin + s < t_end) {
            int r, x0, y0;
            tile_origin(t_begin + s, r, x0, y0);
            issue_tile<1>(A.fix + (int64_t)r * HW, A.edges + (int64_t)r * HW, H, W, x0, y0, S.img[s], S.edg[s]);
        }
        cp_async_commit();
    }
    for (int t = t_begin; t < t_end; ++t) {
        const int stage = (t - t_begin) % kStages;
        {
            const int tn = t + kStages - 1;          // refill the stage consumed in the previous iteration
            if (tn < t_end) {
                int r, x0, y0;
                tile_origin(tn, r, x0, y0);
                const int sn = (tn - t_begin) % kStages;
                issue_tile<1>(A.fix + (int64_t)r * HW, A.edges + (int64_t)r * HW, H, W, x0, y0, S.img[sn], S.edg[sn]);
            }
            cp_async_commit();
        }
        cp_async_wait<kStages - 1>();
        __syncthreads();
        int r, x0, y0;
        tile_origin(t, r, x0, y0);
        double* img = S.img[stage];
        for (int k = tid; k < IW * (kFTY + 2); k += kFNT)
            img[k] = (double)(long long)reinterpret_cast<const unsigned long long*>(img)[k] * kFixToIwe;
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ty = threadIdx.y + 8 * h;
            const int x = x0 + threadIdx.x, y = y0 + ty;
            if (x < W && y < H) {
                const double* p = img + (ty + 1) * IW + threadIdx.x + 1;
                double gx, gy;
                scharr_at(p, IW, gx, gy);
                const double I = *p;
                const double E = S.edg[stage][ty * kFTX + threadIdx.x];
                A.iwe[(int64_t)r * HW + (int64_t)y * W + x] = I;
                FusedAcc o;
                o.sq = gx * gx + gy * gy; o.sI = I; o.sI2 = I * I; o.sEI = E * I; o.mn = I; o.cmn = 1.0; o.mx = I; o.cmx = 1.0;
                acc.merge(o);
            }
        }
        const bool flush = (t + 1 == t_end) || ((t + 1) / tpi != r);
        if (flush) {
            const FusedAcc a = fused_block_reduce(acc, S.red);
            if (tid == 0) {
                double* d = A.part + ((int64_t)r * G + b) * kFPart;
                d[0] = a.sq; d[1] = a.sI; d[2] = a.sI2; d[3] = a.sEI; d[4] = a.mn; d[5] = a.cmn; d[6] = a.mx; d[7] = a.cmx;
            }
            acc.init();
        }
        __syncthreads();                             // stage may be refilled by the next iteration
    }
    cp_async_wait<0>();
    __threadfence();
    cooperative_groups::this_grid().sync();

    // phase 3 prologue first (its loads fly while phase 2 reduces): the float64 images are complete after the barrier
    if (A.want_grad) {
#pragma unroll
        for (int s = 0; s < kStages - 1; ++s) {
            if (t_begin + s < t_end) {
                int r, x0, y0;
                tile_origin(t_begin + s, r, x0, y0);
                issue_tile<2>(A.iwe + (int64_t)r * HW, A.edges + (int64_t)r * HW, H, W, x0, y0, S.img[s], S.edg[s]);
            }
            cp_async_commit();
        }
    }

    // ---- phase 2: global statistics of the reference images this CTA owns (CTA 0: all, plus the loss) ----------------------
    {
        if (tid < R) {                               // cotangent scales from the zero-warp constants (losses.py:176-177)
            const double w = A.sc->weights[tid];
            const double a_r = -A.alpha * w / ((A.sc->zero[0].contrast + kEps) * R);
            const double b_r = A.beta * w / ((-A.sc->zero[tid].mse + kEps) * R);
            S.coefA[tid] = a_r * (2.0 / (double)HW);
            S.coefB[tid] = b_r * (-2.0 / (double)HW);
        }
        __syncthreads();
        const int q_lo = (b == 0) ? 0 : r_first, q_hi = (b == 0) ? R - 1 : r_last;
        for (int q = q_lo; q <= q_hi; ++q) {
            FusedAcc a;
            a.init();
            for (int k = tid; k < G; k += kFNT) {
                const double* d = A.part + ((int64_t)q * G + k) * kFPart;
                FusedAcc o;
                o.sq = __ldcg(d + 0); o.sI = __ldcg(d + 1); o.sI2 = __ldcg(d + 2); o.sEI = __ldcg(d + 3);
                o.mn = __ldcg(d + 4); o.cmn = __ldcg(d + 5); o.mx = __ldcg(d + 6); o.cmx = __ldcg(d + 7);
                a.merge(o);
            }
            a = fused_block_reduce(a, S.red);
            if (tid == 0) S.st[q] = fused_stats(a, (double)HW, A.sc->sumE[q], A.sc->sumE2[q], S.coefB[q]);
        }
        __syncthreads();
        if (b == 0 && tid == 0) {
            for (int q = 0; q < R; ++q) { A.sc->ref[q] = S.st[q]; A.sc->coefA[q] = S.coefA[q]; A.sc->coefB[q] = S.coefB[q]; A.sc->coefD[q] = 0.0; }
            scalars_loss(A.sc, R, A.alpha, A.beta, A.gamma, 0.0, A.use_tv, 0, A.loss_out);
        }
    }

    // ---- phase 3: clear the fixed-point cells; d loss / d IWE ----------------------------------------------------------------
    for (int t = t_begin; t < t_end; ++t) {
        int r, x0, y0;
        tile_origin(t, r, x0, y0);
        unsigned long long* Fr = A.fix + (int64_t)r * HW;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int x = x0 + threadIdx.x, y = y0 + threadIdx.y + 8 * h;
            if (x < W && y < H) Fr[(int64_t)y * W + x] = 0ull;
        }
        if (!A.want_grad) continue;
        const int stage = (t - t_begin) % kStages;
        {
            const int tn = t + kStages - 1;
            if (tn < t_end) {
                int rn, xn, yn;
                tile_origin(tn, rn, xn, yn);
                const int sn = (tn - t_begin) % kStages;
                issue_tile<2>(A.iwe + (int64_t)rn * HW, A.edges + (int64_t)rn * HW, H, W, xn, yn, S.img[sn], S.edg[sn]);
            }
            cp_async_commit();
        }
        cp_async_wait<kStages - 1>();
        __syncthreads();
        const double* tile = S.img[stage];
        for (int k = tid; k < GW * GH; k += kFNT) {
            const int ly = k / GW, lx = k - ly * GW;
            const int y = y0 + ly - 1, x = x0 + lx - 1;
            double gx = 0.0, gy = 0.0;
            if (x >= 0 && x < W && y >= 0 && y < H) scharr_at(tile + (ly + 1) * PW + lx + 1, PW, gx, gy);
            S.gxs[k] = gx; S.gys[k] = gy;
        }
        __syncthreads();
        const Stats st = S.st[r];
        const double cA = S.coefA[r], cB = S.coefB[r];
        const double g_M = -st.s2 / (st.D * st.D);
        const double g_m = -st.s1 / st.D + st.s2 / (st.D * st.D);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ty = threadIdx.y + 8 * h;
            const int x = x0 + threadIdx.x, y = y0 + ty;
            if (x >= W || y >= H) continue;
            const double I = tile[(ty + 2) * PW + threadIdx.x + 2];
            const int64_t p = (int64_t)y * W + x;
            const double adj = scharr_adjoint_at(S.gxs + (ty + 1) * GW + threadIdx.x + 1, S.gys + (ty + 1) * GW + threadIdx.x + 1, GW);
            const double c = I - st.mn;
            const double gN = cB * (S.edg[stage][ty * kFTX + threadIdx.x] - c / st.D);
            double out = cA * adj + gN / st.D;
            if (I == st.mn) out += g_m / st.cnt_min;
            if (I == st.mx) out += g_M / st.cnt_max;
            if (A.dldi != nullptr) A.dldi[r * HW + p] = out;
            A.dldi32[r * HW + p] = (float)(out * kInv2Pi);
        }
        __syncthreads();                             // stage / gxs / gys may be rewritten by the next iteration
    }
    cp_async_wait<0>();
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

}  // namespace eincm
