// Edge images of a window from its grayscale frames (the step before the hot path).
//
// The reference stages, per reference time (src/experiments/e00/exp_mgr.py:343-350),
//     edges_r = normalize_to_unit_range(smoothen_edges(image_to_edge(u8 frame)))
// with OpenCV / SciPy on the CPU:
//   image_to_edge                 = cv.Canny(img, th1, th2, None, 3, L2gradient=True)       src/utils/img_utils.py:194-211
//   smoothen_edges                = cv.GaussianBlur(float64 edge image, ksize from sigma)   src/utils/img_utils.py:213-222
//   eincm_inv_exp_dist_transform  = 1 - normalize(1 - exp(-EDT(~edge) / alpha))             src/utils/img_utils.py:231-235
//   normalize_to_unit_range       = (I - min) / (max - min + eps)                           src/utils/img_utils.py:24-25
// Here all images of a window go through the same launches (blockIdx.z = image).  Canny is integer work and BIT-EXACT with
// cv.Canny: Sobel 3x3 with replicated border in int16, squared magnitude in int32, the TG22 fixed-point sector test, '>' / '>='
// tie rules, hysteresis as connected components (8-neighbourhood) of the candidates that hold a strong candidate - a lock-free
// union-find instead of OpenCV's serial stack (the set of edge pixels does not depend on the visiting order).  The Euclidean
// distance transform is exact (integer squared distances, column scan + row envelope).  The Gaussian / exponential stages are
// float64; their summation order differs from OpenCV's, parity is to 1e-12.
#pragma once
#include "common.cuh"

namespace eincm {

constexpr int kCannyShift = 15;
constexpr int kTG22 = 13573;              // (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5)
constexpr int kEdgeTX = 32, kEdgeTY = 16; // pixels per CTA of the tiled kernels

// ---- Canny stage 1: Sobel, magnitude, non-maximum suppression -> map (0 weak candidate, 1 no edge, 2 strong candidate) and the
// initial union-find forest (label = own index for candidates, -1 otherwise)
__global__ void __launch_bounds__(kEdgeTX* kEdgeTY)
k_canny_nms(const uint8_t* __restrict__ img, int H, int W, int lo, int hi, uint8_t* __restrict__ map, int* __restrict__ label) {
    constexpr int SW = kEdgeTX + 4, SH = kEdgeTY + 4;           // pixels: halo 2
    constexpr int MW = kEdgeTX + 2, MH = kEdgeTY + 2;           // magnitudes: halo 1
    __shared__ uint8_t pix[SH][SW];
    __shared__ int mag[MH][MW];
    __shared__ short sdx[MH][MW], sdy[MH][MW];
    const int64_t off = (int64_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * kEdgeTX, y0 = blockIdx.y * kEdgeTY;
    const int tid = threadIdx.y * kEdgeTX + threadIdx.x;
    for (int i = tid; i < SW * SH; i += kEdgeTX * kEdgeTY) {
        const int sy = i / SW, sx = i - sy * SW;
        const int gy = min(max(y0 + sy - 2, 0), H - 1), gx = min(max(x0 + sx - 2, 0), W - 1);     // BORDER_REPLICATE
        pix[sy][sx] = img[off + (int64_t)gy * W + gx];
    }
    __syncthreads();
    for (int i = tid; i < MW * MH; i += kEdgeTX * kEdgeTY) {
        const int my = i / MW, mx = i - my * MW;
        const int gy = y0 + my - 1, gx = x0 + mx - 1;
        int dx = 0, dy = 0, m = 0;                               // the magnitude buffer has a zero border outside the image
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const int a = pix[my][mx], b = pix[my][mx + 1], c = pix[my][mx + 2];
            const int d = pix[my + 1][mx], f = pix[my + 1][mx + 2];
            const int g = pix[my + 2][mx], h = pix[my + 2][mx + 1], k = pix[my + 2][mx + 2];
            dx = (c - a) + 2 * (f - d) + (k - g);
            dy = (g - a) + 2 * (h - b) + (k - c);
            m = dx * dx + dy * dy;
        }
        sdx[my][mx] = (short)dx; sdy[my][mx] = (short)dy; mag[my][mx] = m;
    }
    __syncthreads();
    const int gx = x0 + threadIdx.x, gy = y0 + threadIdx.y;
    if (gx >= W || gy >= H) return;
    const int mx = threadIdx.x + 1, my = threadIdx.y + 1;
    const int m = mag[my][mx];
    uint8_t out = 1;
    if (m > lo) {
        const int xs = sdx[my][mx], ys = sdy[my][mx];
        const int x = abs(xs), y = abs(ys) << kCannyShift;
        const int tg22x = x * kTG22;
        bool is_max;
        if (y < tg22x) {
            is_max = m > mag[my][mx - 1] && m >= mag[my][mx + 1];
        } else {
            const int tg67x = tg22x + (x << (kCannyShift + 1));
            if (y > tg67x) {
                is_max = m > mag[my - 1][mx] && m >= mag[my + 1][mx];
            } else {
                const int s = (xs ^ ys) < 0 ? -1 : 1;
                is_max = m > mag[my - 1][mx - s] && m > mag[my + 1][mx + s];
            }
        }
        if (is_max) out = m > hi ? 2 : 0;
    }
    const int64_t i = (int64_t)gy * W + gx;
    map[off + i] = out;
    label[off + i] = out != 1 ? (int)i : -1;
}

// ---- union-find over the candidates of one image (labels are pixel indices inside the image; roots satisfy label[i] == i)
__device__ __forceinline__ int uf_find(const int* label, int i) {
    const volatile int* l = label;                               // other threads re-root trees concurrently (labels only decrease)
    int p = l[i];
    while (p != i) { i = p; p = l[i]; }
    return i;
}
__device__ __forceinline__ void uf_union(int* label, int a, int b) {
    for (;;) {
        a = uf_find(label, a);
        b = uf_find(label, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }            // the larger root is attached to the smaller one
        const int old = atomicMin(label + a, b);
        if (old == a) return;
        a = old;                                                 // somebody re-rooted a meanwhile: merge that tree with b
    }
}

// candidates are joined with their W, NW, N, NE candidate neighbours (every 8-neighbour pair is seen once)
__global__ void k_canny_link(const uint8_t* __restrict__ map, int H, int W, int* __restrict__ label_all) {
    const int64_t off = (int64_t)blockIdx.z * H * W;
    int* label = label_all + off;
    const int gx = blockIdx.x * blockDim.x + threadIdx.x, gy = blockIdx.y * blockDim.y + threadIdx.y;
    if (gx >= W || gy >= H) return;
    const int i = gy * W + gx;
    if (map[off + i] == 1) return;
    if (gx > 0 && map[off + i - 1] != 1) uf_union(label, i, i - 1);
    if (gy > 0) {
        const int up = i - W;
        if (map[off + up] != 1) uf_union(label, i, up);
        if (gx > 0 && map[off + up - 1] != 1) uf_union(label, i, up - 1);
        if (gx + 1 < W && map[off + up + 1] != 1) uf_union(label, i, up + 1);
    }
}

// components that hold a strong candidate: flag[root] = 1
__global__ void k_canny_seed(const uint8_t* __restrict__ map, int64_t HW, const int* __restrict__ label_all, uint8_t* __restrict__ flag_all) {
    const int64_t off = (int64_t)blockIdx.y * HW;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x)
        if (map[off + i] == 2) flag_all[off + uf_find(label_all + off, (int)i)] = 1;
}

// edge image: 255 on the candidates of flagged components (cv.Canny's output)
__global__ void k_canny_out(const uint8_t* __restrict__ map, int64_t HW, const int* __restrict__ label_all, const uint8_t* __restrict__ flag_all,
                            uint8_t* __restrict__ edge_all) {
    const int64_t off = (int64_t)blockIdx.y * HW;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
        uint8_t e = 0;
        if (map[off + i] != 1 && flag_all[off + uf_find(label_all + off, (int)i)]) e = 255;
        edge_all[off + i] = e;
    }
}

// ---- min / max of an image through ordered 64-bit keys (atomicMin / atomicMax on unsigned long long)
__device__ __forceinline__ unsigned long long ordered_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_double(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
struct MinMaxKeys { unsigned long long mn, mx; };

__global__ void k_minmax_init(MinMaxKeys* mm, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { mm[i].mn = 0xffffffffffffffffull; mm[i].mx = 0ull; }
}

template <int NT>
__device__ __forceinline__ void block_minmax_to_global(double vmin, double vmax, MinMaxKeys* mm) {
    __shared__ double sh[NT / 32];
    const double bmin = block_reduce<NT>(vmin, OpMin(), sh);
    const double bmax = block_reduce<NT>(vmax, OpMax(), sh);
    if (linear_tid() == 0) {
        atomicMin(&mm->mn, ordered_key(bmin));
        atomicMax(&mm->mx, ordered_key(bmax));
    }
}

// ---- smoothen_edges: separable Gaussian (taps from the host: cv.getGaussianKernel), BORDER_REFLECT_101, float64
constexpr int kGaussMaxTaps = 33;
struct GaussTaps { int n; double w[kGaussMaxTaps]; };

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

__global__ void __launch_bounds__(kEdgeTX* kEdgeTY)
k_edge_gauss(const uint8_t* __restrict__ edge_all, int H, int W, const __grid_constant__ GaussTaps taps, double* __restrict__ out_all, MinMaxKeys* mm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int h = taps.n / 2;
    const int SW = kEdgeTX + 2 * h, SH = kEdgeTY + 2 * h;
    uint8_t* src = smem_raw;                                               // [SH][SW]
    double* rowf = reinterpret_cast<double*>(smem_raw + ((SW * SH + 15) & ~15));   // [SH][kEdgeTX]
    const int64_t off = (int64_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * kEdgeTX, y0 = blockIdx.y * kEdgeTY;
    const int tid = threadIdx.y * kEdgeTX + threadIdx.x;
    for (int i = tid; i < SW * SH; i += kEdgeTX * kEdgeTY) {
        const int sy = i / SW, sx = i - sy * SW;
        src[i] = edge_all[off + (int64_t)reflect101(y0 + sy - h, H) * W + reflect101(x0 + sx - h, W)];
    }
    __syncthreads();
    for (int i = tid; i < SH * kEdgeTX; i += kEdgeTX * kEdgeTY) {
        const int sy = i / kEdgeTX, sx = i - sy * kEdgeTX;
        double acc = 0.0;
        for (int k = 0; k < taps.n; ++k) acc += taps.w[k] * (double)src[sy * SW + sx + k];
        rowf[i] = acc;
    }
    __syncthreads();
    const int gx = x0 + threadIdx.x, gy = y0 + threadIdx.y;
    const bool inside = gx < W && gy < H;
    double acc = 0.0;
    for (int k = 0; k < taps.n; ++k) acc += taps.w[k] * rowf[(threadIdx.y + k) * kEdgeTX + threadIdx.x];
    if (inside) out_all[off + (int64_t)gy * W + gx] = acc;
    block_minmax_to_global<kEdgeTX * kEdgeTY>(inside ? acc : INFINITY, inside ? acc : -INFINITY, mm + blockIdx.z);
}

// ---- eincm_inv_exp_dist_transform: exact Euclidean distance to the nearest edge pixel
constexpr int kEdtInf = 1 << 29;           // "no edge pixel in this column" (squared distances use 64 bits)

// vertical distance to the nearest edge pixel of the column (one thread per column and image)
__global__ void k_edt_columns(const uint8_t* __restrict__ edge_all, int H, int W, int* __restrict__ g_all) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const int64_t off = (int64_t)blockIdx.y * H * W;
    int run = kEdtInf;
    for (int y = 0; y < H; ++y) {
        run = edge_all[off + (int64_t)y * W + x] ? 0 : min(run + 1, kEdtInf);
        g_all[off + (int64_t)y * W + x] = run;
    }
    run = kEdtInf;
    for (int y = H - 1; y >= 0; --y) {
        run = edge_all[off + (int64_t)y * W + x] ? 0 : min(run + 1, kEdtInf);
        const int64_t i = off + (int64_t)y * W + x;
        g_all[i] = min(g_all[i], run);
    }
}

// one CTA per row: d^2(x) = min over x' of (x - x')^2 + g(x')^2, searched outwards from x until (x - x')^2 reaches the best value;
// writes 1 - exp(-d / alpha) and its min / max
__global__ void __launch_bounds__(256)
k_edt_rows(const int* __restrict__ g_all, int H, int W, double alpha, double* __restrict__ out_all, MinMaxKeys* mm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    long long* g2 = reinterpret_cast<long long*>(smem_raw);                // [W]
    const int y = blockIdx.x;
    const int64_t off = (int64_t)blockIdx.y * H * W + (int64_t)y * W;
    for (int x = threadIdx.x; x < W; x += 256) { const long long g = g_all[off + x]; g2[x] = g * g; }
    __syncthreads();
    double vmin = INFINITY, vmax = -INFINITY;
    for (int x = threadIdx.x; x < W; x += 256) {
        long long best = g2[x];
        for (int d = 1; d < W; ++d) {
            const long long dd = (long long)d * d;
            if (dd >= best) break;
            if (x - d >= 0) best = min(best, dd + g2[x - d]);
            if (x + d < W) best = min(best, dd + g2[x + d]);
        }
        const double dist = sqrt((double)best);
        const double e = 1.0 - exp(-dist / alpha);
        out_all[off + x] = e;
        vmin = fmin(vmin, e); vmax = fmax(vmax, e);
    }
    block_minmax_to_global<256>(vmin, vmax, mm + blockIdx.y);
}

// ---- normalize_to_unit_range in place; INVERT: 1 - normalised value (the last step of eincm_inv_exp_dist_transform, followed by
// the normalisation of exp_mgr.py:344, which maps [0, 1] data onto itself up to the eps in the denominator)
template <bool IEDT>
__global__ void k_edge_normalize(double* __restrict__ img_all, int64_t HW, const MinMaxKeys* __restrict__ mm, bool outer) {
    const int64_t off = (int64_t)blockIdx.y * HW;
    const double mn = key_to_double(mm[blockIdx.y].mn), mx = key_to_double(mm[blockIdx.y].mx);
    const double den = (mx - mn) + kEps;
    // IEDT: v = 1 - (e - mn) / den lies in [1 - (mx - mn) / den, 1]; the outer normalize_to_unit_range uses ITS min / max
    const double lo2 = 1.0 - (mx - mn) / den, hi2 = 1.0 - (mn - mn) / den;
    const double den2 = (hi2 - lo2) + kEps;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = img_all[off + i];
        if (IEDT) {
            const double inv = 1.0 - (v - mn) / den;
            img_all[off + i] = outer ? (inv - lo2) / den2 : inv;
        } else {
            img_all[off + i] = (v - mn) / den;
        }
    }
}


// ---- non-local-means denoise of the frames before Canny (preprocess_image, src/utils/img_utils.py:147-157:
// cv.fastNlMeansDenoising(img, None, h, template_win_size, search_win_size), uint8, one channel) -----------------------------
// OpenCV's algorithm (modules/photo/src/fast_nlmeans_denoising_invoker.hpp) is integer and table-driven, restated bit-exactly:
// for every pixel and every offset of the search window, dist = sum over the template window of squared differences between the
// patch around the candidate and the patch around the pixel (BORDER_REFLECT_101); weight = table[dist >> shift] (fixed point,
// table built on the host exactly like OpenCV builds it); output = (sum weight * candidate + sum weight / 2) / sum weight.
// One thread per pixel, the frame tile with its halo (search / 2 + template / 2) in shared memory.
constexpr int kNlmTX = 32, kNlmTY = 8;

__global__ void __launch_bounds__(kNlmTX* kNlmTY)
k_nlm_denoise(const uint8_t* __restrict__ img, int H, int W, int th /* template half */, int sh /* search half */, int shift,
              const int* __restrict__ weight_table, uint8_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = th + sh, SW = kNlmTX + 2 * b, SH = kNlmTY + 2 * b;
    uint8_t* tile = smem_raw;                                                // [SH][SW]
    const int64_t off = (int64_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * kNlmTX, y0 = blockIdx.y * kNlmTY;
    const int tid = threadIdx.y * kNlmTX + threadIdx.x;
    for (int i = tid; i < SW * SH; i += kNlmTX * kNlmTY) {
        const int sy = i / SW, sx = i - sy * SW;
        tile[i] = img[off + (int64_t)reflect101(y0 + sy - b, H) * W + reflect101(x0 + sx - b, W)];
    }
    __syncthreads();
    const int gx = x0 + threadIdx.x, gy = y0 + threadIdx.y;
    if (gx >= W || gy >= H) return;
    const uint8_t* me = tile + (threadIdx.y + b) * SW + threadIdx.x + b;    // the pixel inside the tile
    unsigned int est = 0u, wsum = 0u;                                        // OpenCV sizes the fixed point so that both fit 31 bits
    for (int y = -sh; y <= sh; ++y) {
        for (int x = -sh; x <= sh; ++x) {
            const uint8_t* cand = me + y * SW + x;
            int dist = 0;
            for (int ty = -th; ty <= th; ++ty)
                for (int tx = -th; tx <= th; ++tx) {
                    const int d = (int)cand[ty * SW + tx] - (int)me[ty * SW + tx];
                    dist += d * d;
                }
            const unsigned int wgt = (unsigned int)__ldg(weight_table + (dist >> shift));
            est += wgt * (unsigned int)cand[0];
            wsum += wgt;
        }
    }
    const unsigned int v = (est + wsum / 2u) / wsum;                          // the pixel itself always has the full weight: wsum > 0
    out[off + (int64_t)gy * W + gx] = (uint8_t)min(v, 255u);
}

}  // namespace eincm
