// Fused image-space pass of the default (tile-privatised splat, delta == 0) path: two kernels per evaluation
// instead of compose + A + scalars + B + scalars + C.
//
//   k_image_stats: fixed-point image -> float64 image of warped events, Scharr contrast, min / max with
//                 tie counts, and the moments  sum I, sum I^2, sum E*I  from which the min-max-normalised MSE and the sums of
//                 its backward follow algebraically once the GLOBAL min / max are known:
//                     N = (I - m)/D,  D = max - m + eps
//                     sum (E-N)^2     = sum E^2 - 2/D (sum EI - m sum E) + Q/D^2,      Q = sum I^2 - 2 m sum I + HW m^2
//                     s1 = sum gN     = cb (sum E - (sum I - HW m)/D)                   gN = cb (E - N)
//                     s2 = sum gN(I-m)= cb ((sum EI - m sum E) - Q/D)
//                 so no second pass over the images is needed for the statistics (reference: src/utils/img_utils.py:24-25,
//                 src/eincm/objectives/correlation_objectives.py:25-26, contrast_objectives.py:22-25, src/eincm/losses.py:171-193);
//                 It also applies the adjoint of the Scharr pair to (Gx, Gy) - the part of d loss / d IWE that does not depend on
//                 the global statistics - and its last CTA reduces the per-CTA partials and evaluates the loss.
//   k_image_grad:  pointwise  d loss / d IWE = cA * adjoint + gN / D + min/max tie terms  (float64 + float32/2pi copies, as
//                 k_img_C), and clears the fixed-point images for the next evaluation.
#pragma once

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

// ---- k_image_stats ---------------------------------------------------------------------------------------------------------
// Row-band formulation: a CTA owns a contiguous run of complete image rows (R*H rows laid end to end are split evenly over the
// grid) and handles it in sub-bands of <= band_rows rows.  A sub-band (plus two halo rows on either side and a zero column on
// either side) is brought into shared memory with one burst of asynchronous copies - all loads of the CTA in flight at once -
// and then processed without further global reads: a thread owns the same CPT columns (tid, tid + 256, ...) in every row, so
// the 3x3 stencils need no index arithmetic, global traffic is perfectly coalesced row segments, and every thread accumulates
// its statistics over all its pixels before the single block reduction per (CTA, reference image).  The last CTA to finish
// (ticket) reduces the per-CTA partials of every reference image in a fixed order (deterministic) and evaluates the loss;
// the kernel boundary before k_image_grad replaces the grid-wide barrier of a cooperative formulation (measured: launch of a
// cooperative kernel + barrier + redundant second-level reduction cost more than the second launch).
constexpr int kCoopMaxRefs = EINCM_MAX_REFS;
constexpr int kBandNT = 256;
constexpr int kMaxCPT = 6;                          // columns per thread (template parameter 1..6): sensors up to 1536 px wide
constexpr int kBandRowsMax = 8;                     // rows of a sub-band (chosen by the host: ImageStatsArgs::band_rows)
constexpr int kStatsTicket = 6;                     // DevScalars::counters slot of the "last CTA" ticket

// One cell of the image pass as the fused backward fill reads it (k_backward_fold): ONE 16-byte load per window cell instead of
// three loads from three images.
struct __align__(16) CellRec { double I; float adj; float e; };   // image value, Scharr adjoint (up to coefA), float32 edge value

struct ImageStatsArgs {
    CellRec* rec;                   // non-null: write [R][H*W] cell records instead of the dense iwe / adj32 images
    const float* e32;               // [R][H*W] float32 edge images (only read when rec != null)
    const unsigned long long* fix;  // [R][H*W] fixed-point images of warped events
    const double* edges;            // [R][H*W]
    double* iwe;                    // [R][H*W] float64 images (out)
    float* adj32;                   // [R][H*W] adjoint of the Scharr pair applied to (Gx, Gy) (out; d contrast / d IWE up to cA)
    double* part;                   // [R][grid][kFPart]
    DevScalars* sc;
    double* loss_out;
    double* zero_buf;               // buffers cleared for the event backward pass (dense flow-field gradient, theta gradient) or null
    double* zero_buf2;
    int n_zero, n_zero2;
    int H, W, R;
    double alpha, beta, gamma;
    int use_tv;
    int band_rows;                  // rows of a sub-band held in shared memory (<= kBandRowsMax; shared memory is sized for it)
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

// 8-byte asynchronous global -> shared copy; `valid == false` zero-fills (rows outside the image)
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

struct BandSmemTail {
    double red[kBandNT / 32][kFPart];
    Stats st[kCoopMaxRefs];
    double coefA[kCoopMaxRefs], coefB[kCoopMaxRefs];
    double wts[kCoopMaxRefs], zero_mse[kCoopMaxRefs], sumE[kCoopMaxRefs], sumE2[kCoopMaxRefs], zero_contrast;
    int is_last;
};

// image rows (halo 2) + Scharr pair of rows (halo 1, float32) of a sub-band of B rows
__host__ __device__ inline size_t image_pass_smem_bytes(int W, int B) {
    return (size_t)((B + 4) + (B + 2)) * (W + 2) * sizeof(double) + sizeof(BandSmemTail);
}

template <int CPT>
__global__ void __launch_bounds__(kBandNT, 2)
k_image_stats(const ImageStatsArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int H = A.H, W = A.W, R = A.R, Wp = W + 2;
    const int B = A.band_rows;
    double* bandI = reinterpret_cast<double*>(smem_raw);             // [B + 4][Wp]: image rows sa-2 .. sa+B+1
    float* gxs = reinterpret_cast<float*>(bandI + (B + 4) * Wp);     // [B + 2][Wp] Scharr x of rows sa-1 .. sa+B
    float* gys = gxs + (B + 2) * Wp;                                 // [B + 2][Wp] Scharr y
    BandSmemTail& S = *reinterpret_cast<BandSmemTail*>(gys + (B + 2) * Wp);
    const int HW = H * W;
    const int tid = threadIdx.x;
    const int G = gridDim.x, b = blockIdx.x;
    const int RH = R * H;
    const int row_begin = (int)(((long long)RH * b) / G), row_end = (int)(((long long)RH * (b + 1)) / G);
    const int r_first = row_begin < row_end ? row_begin / H : 0, r_last = row_begin < row_end ? (row_end - 1) / H : -1;

    // programmatic dependent launch: the next kernel of the evaluation may be scheduled as soon as every CTA of this one is resident;
    // this kernel itself may have been scheduled while the splat was still running - nothing but shared memory and per-window
    // constants is touched before the wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // zero columns (never written afterwards), identity partials, accumulators of the event backward pass
    for (int k = tid; k < B + 4; k += kBandNT) { bandI[k * Wp] = 0.0; bandI[k * Wp + W + 1] = 0.0; }
    for (int k = tid; k < 2 * (B + 2); k += kBandNT) { gxs[k * Wp] = 0.f; gxs[k * Wp + W + 1] = 0.f; }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int k = tid; k < R * kFPart; k += kBandNT) {
        const int q = k / kFPart, f = k % kFPart;
        A.part[(q * G + b) * kFPart + f] = (f == 4) ? INFINITY : ((f == 6) ? -INFINITY : 0.0);
    }
    if (A.zero_buf != nullptr)
        for (int k = b * kBandNT + tid; k < A.n_zero; k += G * kBandNT) A.zero_buf[k] = 0.0;
    if (A.zero_buf2 != nullptr)
        for (int k = b * kBandNT + tid; k < A.n_zero2; k += G * kBandNT) A.zero_buf2[k] = 0.0;
    // per-window constants of the tail (cotangent scales from the zero-warp image, losses.py:176-177): fetched now, their
    // latency hides behind the band loop
    if (tid < R) {
        const double w = A.sc->weights[tid], zc = A.sc->zero[0].contrast, zm = A.sc->zero[tid].mse;
        const double a_r = -A.alpha * w / ((zc + kEps) * R);
        const double b_r = A.beta * w / ((-zm + kEps) * R);
        S.coefA[tid] = a_r * (2.0 / (double)HW);
        S.coefB[tid] = b_r * (-2.0 / (double)HW);
        S.wts[tid] = w; S.zero_mse[tid] = zm; S.sumE[tid] = A.sc->sumE[tid]; S.sumE2[tid] = A.sc->sumE2[tid];
        if (tid == 0) S.zero_contrast = zc;
    }

    for (int r = r_first; r <= r_last; ++r) {
        const int ya = max(row_begin - r * H, 0), yb = min(row_end - r * H, H);
        const unsigned long long* Fr = A.fix + r * HW;
        const double* Er = A.edges + r * HW;
        double* Ir = A.iwe + r * HW;
        float* Ar = A.adj32 + r * HW;
        CellRec* Rr = A.rec != nullptr ? A.rec + r * HW : nullptr;
        const float* E32r = A.e32 + r * HW;
        FusedAcc acc;
        acc.init();
        int cnt_mn = 0, cnt_mx = 0;                   // tie counts of the running min / max (integers: branch-free update)
        for (int sa = ya; sa < yb; sa += B) {
            const int sb = min(sa + B, yb);
            __syncthreads();                           // previous readers of the band are done
            for (int y = sa - 2; y < sb + 2; ++y) {    // fixed-point rows sa-2 .. sb+1 -> bandI, row y at slot y - sa + 2
                double* dst = bandI + (y - sa + 2) * Wp + 1;
                const bool in = y >= 0 && y < H;
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const int x = tid + c * kBandNT;
                    if (x < W) cp_async8(dst + x, in ? (const void*)(Fr + y * W + x) : (const void*)Fr, in);
                }
            }
            double e_next[CPT];                        // edge row of the next own row: register prefetch, one row ahead
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int x = tid + c * kBandNT;
                e_next[c] = (x < W) ? __ldg(Er + sa * W + x) : 0.0;
            }
            cp_async_commit_wait_all();
            for (int y = sa - 2; y < sb + 2; ++y) {    // fixed point -> float64, each thread the cells it copied
                double* row = bandI + (y - sa + 2) * Wp + 1;
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const int x = tid + c * kBandNT;
                    // sums are far below 2^52: (2^52 | v) reinterpreted as float64 is exactly 2^52 + v
                    if (x < W) row[x] = __dsub_rn(__longlong_as_double(0x4330000000000000ll | reinterpret_cast<const long long*>(row)[x]), 4503599627370496.0) * kFixToIwe;
                }
            }
            __syncthreads();
            // Scharr pair of rows sa-1 .. sb (zero outside the image: the 'same' output only exists inside); statistics of the
            // rows this CTA owns; float32 copies for the adjoint
            for (int q = sa - 1; q <= sb; ++q) {
                const double* mid = bandI + (q - sa + 2) * Wp + 1;
                const double* up = mid - Wp;
                const double* dn = mid + Wp;
                float* gxr = gxs + (q - sa + 1) * Wp + 1;
                float* gyr = gys + (q - sa + 1) * Wp + 1;
                const bool inside = q >= 0 && q < H, own = q >= sa && q < sb;
                double er[CPT];
#pragma unroll
                for (int c = 0; c < CPT; ++c) er[c] = e_next[c];
                if (own && q + 1 < sb) {
#pragma unroll
                    for (int c = 0; c < CPT; ++c) {
                        const int x = tid + c * kBandNT;
                        e_next[c] = (x < W) ? __ldg(Er + (q + 1) * W + x) : 0.0;
                    }
                }
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const int x = tid + c * kBandNT;
                    if (x < W) {
                        double gx = 0.0, gy = 0.0;
                        if (inside) scharr_vals(dn[x + 1], dn[x - 1], mid[x + 1], mid[x - 1], up[x + 1], up[x - 1], dn[x], up[x], gx, gy);
                        gxr[x] = (float)gx; gyr[x] = (float)gy;
                        if (own) {
                            const double I = mid[x];
                            if (Rr != nullptr) Rr[q * W + x].I = I; else Ir[q * W + x] = I;
                            acc.sq += gx * gx + gy * gy; acc.sI += I; acc.sI2 += I * I; acc.sEI += er[c] * I;
                            cnt_mn = (I < acc.mn) ? 1 : cnt_mn + (I == acc.mn ? 1 : 0);
                            cnt_mx = (I > acc.mx) ? 1 : cnt_mx + (I == acc.mx ? 1 : 0);
                            acc.mn = fmin(acc.mn, I);
                            acc.mx = fmax(acc.mx, I);
                        }
                    }
                }
            }
            __syncthreads();
            for (int y = sa; y < sb; ++y) {
                const float* xm = gxs + (y - sa + 1) * Wp + 1;
                const float* ym = gys + (y - sa + 1) * Wp + 1;
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const int x = tid + c * kBandNT;
                    if (x < W) {
                        const float a = scharr_adjoint_rows_f32(xm - Wp + x, xm + x, xm + Wp + x, ym - Wp + x, ym + Wp + x);
                        if (Rr != nullptr) *reinterpret_cast<float2*>(&Rr[y * W + x].adj) = make_float2(a, __ldg(E32r + y * W + x));
                        else Ar[y * W + x] = a;
                    }
                }
            }
        }
        acc.cmn = (double)cnt_mn; acc.cmx = (double)cnt_mx;
        const FusedAcc a = fused_block_reduce(acc, S.red);
        if (tid == 0) {
            double* d = A.part + (r * G + b) * kFPart;
            d[0] = a.sq; d[1] = a.sI; d[2] = a.sI2; d[3] = a.sEI; d[4] = a.mn; d[5] = a.cmn; d[6] = a.mx; d[7] = a.cmx;
        }
    }

    // ---- tail: the last CTA reduces the partials of every reference image (one WARP per image, fixed order) -----------------
    __threadfence();
    __syncthreads();
    if (tid == 0) S.is_last = atomicAdd(&A.sc->counters[kStatsTicket], 1u) == (unsigned)(G - 1);
    __syncthreads();
    if (!S.is_last) return;
    __threadfence();
    {
        const int lane = tid & 31;
        for (int q = tid >> 5; q < R; q += kBandNT / 32) {
            // only the CTAs whose rows intersect image q hold a non-identity partial: a contiguous range of CTAs.  Four records
            // per lane and round, all loads of a round issued before the first merge (one L2 round trip per round)
            const int k_lo = (int)(((long long)q * H * G) / RH);
            const int k_hi = min(G, (int)((((long long)(q + 1) * H * G) + RH - 1) / RH) + 1);
            FusedAcc a;
            a.init();
            for (int k0 = k_lo + lane; k0 < k_hi; k0 += 4 * 32) {
                double v[4][kFPart];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = min(k0 + 32 * u, k_hi - 1);
                    const double* d = A.part + (q * G + k) * kFPart;
#pragma unroll
                    for (int f = 0; f < kFPart; ++f) v[u][f] = __ldcg(d + f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (k0 + 32 * u < k_hi) {
                        FusedAcc o;
                        o.sq = v[u][0]; o.sI = v[u][1]; o.sI2 = v[u][2]; o.sEI = v[u][3];
                        o.mn = v[u][4]; o.cmn = v[u][5]; o.mx = v[u][6]; o.cmx = v[u][7];
                        a.merge(o);
                    }
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) a.merge(a.shfl_xor(o));
            if (lane == 0) S.st[q] = fused_stats(a, (double)HW, S.sumE[q], S.sumE2[q], S.coefB[q]);
        }
        __syncthreads();
        if (tid == 0) {
            for (int q = 0; q < R; ++q) {
                A.sc->ref[q] = S.st[q]; A.sc->coefA[q] = S.coefA[q]; A.sc->coefB[q] = S.coefB[q]; A.sc->coefD[q] = 0.0;
                // coefficients of the per-cell cotangent for the fused backward fill (same terms as k_image_grad, scaled by 1 / 2 pi)
                const Stats& st = S.st[q];
                const double iD = 1.0 / st.D;
                const double g_M = -st.s2 / (st.D * st.D);
                const double g_m = -st.s1 / st.D + st.s2 / (st.D * st.D);
                CotCoef c;
                c.cA = S.coefA[q] * kInv2Pi;
                c.a1 = S.coefB[q] * iD * kInv2Pi;
                c.a2 = S.coefB[q] * iD * iD * kInv2Pi;
                c.a3 = c.a2 * st.mn;
                c.mn = st.mn; c.mx = st.mx;
                c.tm = g_m / st.cnt_min * kInv2Pi; c.tM = g_M / st.cnt_max * kInv2Pi;
                A.sc->cot[q] = c;
            }
            // final loss (reference src/eincm/losses.py:171-193)
            double s_corr = 0.0, s_con = 0.0;
            for (int q = 0; q < R; ++q) {
                s_corr += (S.wts[q] * (-S.st[q].mse)) / ((-S.zero_mse[q]) + kEps);          // losses.py:176
                s_con += (S.wts[q] * S.st[q].contrast) / (S.zero_contrast + kEps);          // losses.py:177
            }
            const double mean_rel_corr = s_corr / R, mean_rel_contrast = s_con / R;
            const double tv = A.use_tv ? A.sc->tv_sum / (A.sc->tv_cnt + kEps) : 0.0;        // regularizers.py:31-36, losses.py:171
            const double loss = (A.alpha * (mean_rel_contrast * (-1.0)) + A.beta * (mean_rel_corr * (-1.0))) + (A.gamma * tv + 0.0);
            A.sc->loss = loss; A.sc->mean_rel_corr = mean_rel_corr; A.sc->mean_rel_contrast = mean_rel_contrast;
            A.sc->mean_rel_div = 0.0; A.sc->tv = tv;
            if (A.loss_out != nullptr) *A.loss_out = loss;
            A.sc->counters[kStatsTicket] = 0u;
        }
    }
}

// ---- k_image_grad ----------------------------------------------------------------------------------------------------------
// d loss / d IWE_r = cA_r * adjoint(Gx, Gy) + gN / D + (tie-split cotangents of min and max)   with   gN = cB_r (E - (I - m) / D)
// (reverse mode of contrast_objectives.py:22-25, img_utils.py:24-25, correlation_objectives.py:25-26; min / max cotangents
// split evenly among ties like jnp.min / jnp.max).  Pointwise: four cells per thread and round, loads batched.  Also clears the
// fixed-point images (every reader has finished: kernel boundary).
struct ImageGradArgs {
    unsigned long long* fix;        // [R][H*W] cleared (null: left alone)
    const double* edges;
    const double* iwe;
    const float* adj32;
    const DevScalars* sc;
    double* dldi;                   // [R][H*W] float64 d loss / d IWE (debug tap) or null
    float* dldi32;                  // [R][H*W] float32 d loss / d IWE / (2 pi) or null
    int HW, R;
    int want_grad;                  // 0: only clear the fixed-point images
};

__global__ void __launch_bounds__(256)
k_image_grad(const ImageGradArgs A) {
    // programmatic dependent launch (no-ops when launched plainly): let the next kernel be scheduled, wait for the image statistics
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __shared__ double s_mn[kCoopMaxRefs], s_mx[kCoopMaxRefs], s_iD[kCoopMaxRefs], s_cA[kCoopMaxRefs], s_cB[kCoopMaxRefs], s_tm[kCoopMaxRefs], s_tM[kCoopMaxRefs];
    if (threadIdx.x < A.R && A.want_grad) {
        const Stats st = A.sc->ref[threadIdx.x];
        s_mn[threadIdx.x] = st.mn; s_mx[threadIdx.x] = st.mx; s_iD[threadIdx.x] = 1.0 / st.D;
        s_cA[threadIdx.x] = A.sc->coefA[threadIdx.x]; s_cB[threadIdx.x] = A.sc->coefB[threadIdx.x];
        const double g_M = -st.s2 / (st.D * st.D);
        const double g_m = -st.s1 / st.D + st.s2 / (st.D * st.D);
        s_tm[threadIdx.x] = g_m / st.cnt_min; s_tM[threadIdx.x] = g_M / st.cnt_max;
    }
    __syncthreads();
    const int T = gridDim.x * blockDim.x;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < A.R; ++r) {
        // 1 / D is a per-image constant: the two divisions of the reference formula become multiplications (one rounding more)
        const double mn = s_mn[r], mx = s_mx[r], iD = s_iD[r], cA = s_cA[r], cBD = s_cB[r] * s_iD[r], tm = s_tm[r], tM = s_tM[r];
        const double* Er = A.edges + (size_t)r * A.HW;
        const double* Ir = A.iwe + (size_t)r * A.HW;
        const float* Ar = A.adj32 + (size_t)r * A.HW;
        for (int base = gt; base < A.HW; base += 4 * T) {
            double I[4], E[4];
            float a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int p = min(base + u * T, A.HW - 1);
                if (A.want_grad) { I[u] = __ldcg(Ir + p); E[u] = __ldg(Er + p); a[u] = __ldcg(Ar + p); }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int p = base + u * T;
                if (p < A.HW) {
                    if (A.fix != nullptr) A.fix[(size_t)r * A.HW + p] = 0ull;
                    if (A.want_grad) {
                        double out = cA * (double)a[u] + cBD * (E[u] - (I[u] - mn) * iD);
                        if (I[u] == mn) out += tm;
                        if (I[u] == mx) out += tM;
                        if (A.dldi != nullptr) A.dldi[(size_t)r * A.HW + p] = out;
                        if (A.dldi32 != nullptr) A.dldi32[(size_t)r * A.HW + p] = (float)(out * kInv2Pi);
                    }
                }
            }
        }
    }
}

// cell records -> dense float64 image + float32 adjoint (debug taps and the unfused d loss / d IWE kernel, on demand)
__global__ void k_unpack_records(const CellRec* __restrict__ rec, int64_t n, double* __restrict__ iwe, float* __restrict__ adj32) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const CellRec c = rec[i];
        iwe[i] = c.I; adj32[i] = c.adj;
    }
}

// per-window: sum E_r and sum E_r^2 (deterministic single-CTA-per-reference reduction; once per window)
__global__ void __launch_bounds__(1024)
k_edge_sums(const double* __restrict__ edges, int64_t HW, DevScalars* sc, float* __restrict__ e32 /* float32 copy for the fused backward fill, or null */) {
    __shared__ double sh[32];
    const int r = blockIdx.x;
    double s = 0.0, s2 = 0.0;
    for (int64_t p = threadIdx.x; p < HW; p += blockDim.x) {
        const double e = edges[r * HW + p];
        s += e; s2 += e * e;
        if (e32 != nullptr) e32[r * HW + p] = (float)e;
    }
    s = block_reduce<1024>(s, OpSum(), sh);
    s2 = block_reduce<1024>(s2, OpSum(), sh);
    if (threadIdx.x == 0) { sc->sumE[r] = s; sc->sumE2[r] = s2; }
}

}  // namespace eincm
