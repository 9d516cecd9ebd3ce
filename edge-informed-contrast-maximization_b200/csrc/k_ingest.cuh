// Event ingest before staging: what the reference's DSEC loader and experiment manager do in NumPy /
// jax.numpy between the h5 event stream and the (xs, ys, ts) operands of loss_func.
//   rectify_events   src/dataloaders/dsec_loader.py:145-170   (x, y) -> round(rectify_map[y, x]) as int16, out-of-sensor events dropped
//   time normalise   src/experiments/e00/exp_mgr.py:313-321   ts = (t - start) / (end - start + eps), float64
// Index and byte work: bit-exact.  The compaction keeps the event order (the stream stays sorted by time).
#pragma once
#include "common.cuh"

namespace eincm {

constexpr int kIngestNT = 256, kIngestPerThread = 4, kIngestTile = kIngestNT * kIngestPerThread;

// np.round(float32) is round-half-to-even in float32; .astype('int16') truncates the (already integral) value
__device__ __forceinline__ int rect_coord(float v) { return (int)(short)__float2int_rn(v); }

__device__ __forceinline__ bool rectified(const int16_t* x, const int16_t* y, int64_t i, const float* map, int H, int W, int& rx, int& ry) {
    const int xi = x[i], yi = y[i];
    if (xi < 0 || xi >= W || yi < 0 || yi >= H) { rx = ry = -1; return false; }    // the reference asserts raw events are in-sensor
    const float2 m = reinterpret_cast<const float2*>(map)[(int64_t)yi * W + xi];
    rx = rect_coord(m.x); ry = rect_coord(m.y);
    return rx >= 0 && rx < W && ry >= 0 && ry < H;
}

// pass 1: number of surviving events per tile of kIngestTile events
__global__ void __launch_bounds__(kIngestNT)
k_rectify_count(const int16_t* __restrict__ x, const int16_t* __restrict__ y, int64_t n, const float* __restrict__ map, int H, int W,
                unsigned int* __restrict__ tile_count) {
    const int64_t base = (int64_t)blockIdx.x * kIngestTile;
    int c = 0;
#pragma unroll
    for (int k = 0; k < kIngestPerThread; ++k) {
        const int64_t i = base + (int64_t)threadIdx.x * kIngestPerThread + k;
        int rx, ry;
        if (i < n && rectified(x, y, i, map, H, W, rx, ry)) ++c;
    }
    __shared__ int sh[kIngestNT / 32];
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < kIngestNT / 32; ++w) s += sh[w];
        tile_count[blockIdx.x] = (unsigned)s;
    }
}

// pass 2: exclusive scan of the tile counts (one CTA; 64-bit offsets), total -> *n_out
__global__ void __launch_bounds__(1024)
k_rectify_scan(const unsigned int* __restrict__ tile_count, int64_t n_tiles, long long* __restrict__ tile_offset, long long* __restrict__ n_out) {
    __shared__ long long warp_sum[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t b0 = 0; b0 < n_tiles; b0 += 1024) {
        const int64_t i = b0 + threadIdx.x;
        const long long v = i < n_tiles ? (long long)tile_count[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += u;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            long long w = warp_sum[threadIdx.x], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long u = __shfl_up_sync(0xffffffffu, wi, o);
                if (threadIdx.x >= o) wi += u;
            }
            warp_sum[threadIdx.x] = wi - w;                      // exclusive prefix of the warp sums
        }
        __syncthreads();
        const long long excl = carry + warp_sum[threadIdx.x >> 5] + incl - v;
        if (i < n_tiles) tile_offset[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = carry;
}

// pass 3: rectified coordinates of the surviving events, written in order
__global__ void __launch_bounds__(kIngestNT)
k_rectify_scatter(const int16_t* __restrict__ x, const int16_t* __restrict__ y, const int64_t* __restrict__ t, const uint8_t* __restrict__ p,
                  int64_t n, const float* __restrict__ map, int H, int W, const long long* __restrict__ tile_offset,
                  int16_t* __restrict__ x_out, int16_t* __restrict__ y_out, int64_t* __restrict__ t_out, uint8_t* __restrict__ p_out) {
    __shared__ int warp_excl[kIngestNT / 32];
    const int64_t base = (int64_t)blockIdx.x * kIngestTile;
    int rx[kIngestPerThread], ry[kIngestPerThread];
    bool keep[kIngestPerThread];
    int c = 0;
#pragma unroll
    for (int k = 0; k < kIngestPerThread; ++k) {
        const int64_t i = base + (int64_t)threadIdx.x * kIngestPerThread + k;
        keep[k] = i < n && rectified(x, y, i, map, H, W, rx[k], ry[k]);
        c += keep[k] ? 1 : 0;
    }
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += u;
    }
    if ((threadIdx.x & 31) == 31) warp_excl[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < kIngestNT / 32; ++w) { const int v = warp_excl[w]; warp_excl[w] = s; s += v; }
    }
    __syncthreads();
    long long o = tile_offset[blockIdx.x] + warp_excl[threadIdx.x >> 5] + incl - c;
#pragma unroll
    for (int k = 0; k < kIngestPerThread; ++k) {
        if (!keep[k]) continue;
        const int64_t i = base + (int64_t)threadIdx.x * kIngestPerThread + k;
        x_out[o] = (int16_t)rx[k]; y_out[o] = (int16_t)ry[k];
        if (t_out) t_out[o] = t[i];
        if (p_out) p_out[o] = p[i];
        ++o;
    }
}

// exp_mgr.py:313-321: ts (uint64 microseconds) - start_time (int64) promotes to float64 in jax.numpy / NumPy; the span is
// (end - start) + eps in float64
__global__ void k_normalize_times(const int64_t* __restrict__ t_us, int64_t n, double start, double span, double* __restrict__ ts_out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        ts_out[i] = __ddiv_rn(__dsub_rn((double)(unsigned long long)t_us[i], start), span);
}

}  // namespace eincm
