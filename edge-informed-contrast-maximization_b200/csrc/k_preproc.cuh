// Image pre-processing between the non-local-means denoise and the bilateral filter of preprocess_image (reference
// src/utils/img_utils.py:159-181), uint8 frames, bit-exact with OpenCV 4.x (restated and pinned in oracle/edge_oracle.py):
//   cv.createCLAHE(clipLimit, tileGridSize).apply      k_clahe_lut (one CTA per tile: histogram, clip + redistribution, LUT) + k_clahe_apply
//   cv.GaussianBlur(uint8) + cv.addWeighted ("sharpen")  k_sharpen (separable 8-bit fixed-point kernel, BORDER_REFLECT_101, float32 blend)
// Byte / integer work: bound by launch latency at the frame sizes of the datasets (three 640x480 frames per window).
#pragma once
#include "common.cuh"
#include "k_edges.cuh"

namespace eincm {

// ---- CLAHE (modules/imgproc/src/clahe.cpp) ---------------------------------------------------------------------------------------------
// Tiles of tw x th pixels over the frame, extended to a multiple of the tile grid by BORDER_REFLECT_101 on the right / bottom.
// lut[image][tile row][tile column][256].
__global__ void __launch_bounds__(256)
k_clahe_lut(const uint8_t* __restrict__ img_all, int H, int W, int tiles_x, int tiles_y, int tw, int th, int clip, float lut_scale,
            uint8_t* __restrict__ lut_all) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int wsum[8];
    __shared__ unsigned int s_clipped;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int ti = blockIdx.x, tj = blockIdx.y, im = blockIdx.z;
    const uint8_t* img = img_all + (int64_t)im * H * W;
    hist[tid] = 0u;
    if (tid == 0) s_clipped = 0u;
    __syncthreads();
    const int x0 = ti * tw, y0 = tj * th, area = tw * th;
    for (int i = tid; i < area; i += 256) {
        const int yy = i / tw, xx = i - yy * tw;
        // the extension only pads on the right / bottom: reflect101 of an index beyond the last row / column
        atomicAdd(&hist[img[(int64_t)reflect101(y0 + yy, H) * W + reflect101(x0 + xx, W)]], 1u);
    }
    __syncthreads();
    unsigned int hv = hist[tid];
    if (clip > 0) {
        // clip and redistribute: the excess in equal batches to every bin, the residual one by one to every step-th bin from bin 0
        const unsigned int over = hv > (unsigned)clip ? hv - (unsigned)clip : 0u;
        if (over) atomicAdd(&s_clipped, over);
        __syncthreads();
        const unsigned int clipped = s_clipped;
        hv = min(hv, (unsigned)clip);
        const unsigned int batch = clipped / 256u, residual = clipped - batch * 256u;
        hv += batch;
        if (residual != 0u) {
            const unsigned int step = max(256u / residual, 1u);
            if ((unsigned)tid % step == 0u && (unsigned)tid / step < residual) hv += 1u;
        }
    }
    // inclusive scan over the 256 bins
    unsigned int v = hv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    if (lane == 31) wsum[wid] = v;
    __syncthreads();
    unsigned int before = 0u;
    for (int q = 0; q < wid; ++q) before += wsum[q];
    const unsigned int cum = v + before;
    // saturate_cast<uchar>(sum * lutScale): float32 product, round half to even
    const int r = __float2int_rn(__fmul_rn((float)cum, lut_scale));
    lut_all[(((int64_t)im * tiles_y + tj) * tiles_x + ti) * 256 + tid] = (uint8_t)min(max(r, 0), 255);
}

// bilinear blend of the four neighbouring tile LUTs in float32 (every operation rounded on its own, like the scalar code of OpenCV)
__global__ void __launch_bounds__(256)
k_clahe_apply(const uint8_t* __restrict__ img_all, int H, int W, int tiles_x, int tiles_y, float inv_tw, float inv_th,
              const uint8_t* __restrict__ lut_all, uint8_t* __restrict__ out_all) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), im = blockIdx.z;
    if (x >= W || y >= H) return;
    const int64_t off = (int64_t)im * H * W + (int64_t)y * W + x;
    const uint8_t* lut = lut_all + (int64_t)im * tiles_y * tiles_x * 256;
    const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f), tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
    int tx1 = (int)floorf(txf), ty1 = (int)floorf(tyf);
    const float xa = __fsub_rn(txf, (float)tx1), ya = __fsub_rn(tyf, (float)ty1);
    const float xa1 = __fsub_rn(1.0f, xa), ya1 = __fsub_rn(1.0f, ya);
    const int tx2 = min(tx1 + 1, tiles_x - 1), ty2 = min(ty1 + 1, tiles_y - 1);
    tx1 = max(tx1, 0); ty1 = max(ty1, 0);
    const int v = img_all[off];
    const float l11 = (float)lut[((ty1 * tiles_x) + tx1) * 256 + v], l12 = (float)lut[((ty1 * tiles_x) + tx2) * 256 + v];
    const float l21 = (float)lut[((ty2 * tiles_x) + tx1) * 256 + v], l22 = (float)lut[((ty2 * tiles_x) + tx2) * 256 + v];
    const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa)), bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
    const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
    out_all[off] = (uint8_t)min(max(__float2int_rn(res), 0), 255);
}

// ---- sharpen: cv.GaussianBlur of a uint8 image + cv.addWeighted (img_utils.py:163-178) ------------------------------------------------
// The 8-bit fixed-point Gaussian kernel OpenCV uses for uint8 images (taps sum to 256; built on the host like
// getGaussianKernelFixedPoint_ED), applied along the rows, then along the columns: (sum + 2^15) >> 16, BORDER_REFLECT_101.  Then
// out = saturate(rint(src * alpha + blur * beta + gamma)) in float32.
constexpr int kSharpMaxTaps = 63;
constexpr int kSharpTX = 32, kSharpTY = 16;
struct FixedTaps { int n; int w[kSharpMaxTaps]; };

__global__ void __launch_bounds__(kSharpTX* kSharpTY)
k_sharpen(const uint8_t* __restrict__ img_all, int H, int W, const __grid_constant__ FixedTaps taps, float alpha, float beta, float gamma,
          uint8_t* __restrict__ blur_out_all /* or null */, uint8_t* __restrict__ out_all) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int hw = taps.n / 2;
    const int sw = kSharpTX + 2 * hw, shh = kSharpTY + 2 * hw;
    uint8_t* src = smem_raw;                                                     // [shh][sw]
    unsigned short* row = reinterpret_cast<unsigned short*>(smem_raw + (((size_t)sw * shh + 15) & ~(size_t)15));   // [shh][kSharpTX]: <= 255 * 256
    const int tid = threadIdx.y * kSharpTX + threadIdx.x;
    const int x0 = blockIdx.x * kSharpTX, y0 = blockIdx.y * kSharpTY;
    const int64_t off = (int64_t)blockIdx.z * H * W;
    for (int i = tid; i < sw * shh; i += kSharpTX * kSharpTY) {
        const int sy = i / sw, sx = i - sy * sw;
        src[i] = img_all[off + (int64_t)reflect101(y0 + sy - hw, H) * W + reflect101(x0 + sx - hw, W)];
    }
    __syncthreads();
    for (int i = tid; i < kSharpTX * shh; i += kSharpTX * kSharpTY) {
        const int sy = i / kSharpTX, sx = i - sy * kSharpTX;
        int a = 0;
        for (int k = 0; k < taps.n; ++k) a += taps.w[k] * (int)src[sy * sw + sx + k];
        row[i] = (unsigned short)a;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;
    int a = 0;
    for (int k = 0; k < taps.n; ++k) a += taps.w[k] * (int)row[(threadIdx.y + k) * kSharpTX + threadIdx.x];
    const int blur = min(max((a + (1 << 15)) >> 16, 0), 255);
    const int v = src[(threadIdx.y + hw) * sw + threadIdx.x + hw];
    if (blur_out_all != nullptr) blur_out_all[off + (int64_t)y * W + x] = (uint8_t)blur;
    const float r = __fadd_rn(__fadd_rn(__fmul_rn((float)v, alpha), __fmul_rn((float)blur, beta)), gamma);
    out_all[off + (int64_t)y * W + x] = (uint8_t)min(max(__float2int_rn(r), 0), 255);
}

}  // namespace eincm
