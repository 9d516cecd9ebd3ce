// Per-window staging kernels: validate events, bin them by source pixel (counting sort), event mask.
// Replaces the per-evaluation work the reference redoes inside every loss call for theta-independent
// quantities (reference src/eincm/losses.py:54-55, src/utils/theta_utils.py:66-71).
#pragma once
#include "common.cuh"

namespace eincm {

// Sort key of a source pixel.  Tile-major (TS x TS tiles, row-major inside a tile) so that consecutive sorted
// events share a compact 2-D neighbourhood: their warped destinations then share cache lines / smem tiles.
constexpr int kSortTile = 16;

__host__ __device__ __forceinline__ int sort_key(int x, int y, int W, int tiles_x) {
    const int tx = x / kSortTile, ty = y / kSortTile;
    return ((ty * tiles_x + tx) * kSortTile + (y % kSortTile)) * kSortTile + (x % kSortTile);
}

// counts[key] += 1 ; flags events outside the sensor
__global__ void k_histogram(const int16_t* __restrict__ xs, const int16_t* __restrict__ ys, int64_t n,
                            int H, int W, int tiles_x, unsigned int* __restrict__ counts, int* __restrict__ error_flag) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int x = xs[e], y = ys[e];
        if (x < 0 || x >= W || y < 0 || y >= H) { *error_flag = 1; continue; }
        atomicAdd(&counts[sort_key(x, y, W, tiles_x)], 1u);
    }
}

// Exclusive scan of `counts` (n_keys entries) in three launches: block sums, scan of block sums, add-back.
constexpr int kScanBlock = 1024;

__global__ void k_scan_block_sums(const unsigned int* __restrict__ counts, int n_keys, unsigned int* __restrict__ block_sums) {
    __shared__ unsigned int sh[32];
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    unsigned int v = (i < n_keys) ? counts[i] : 0u;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = sh[threadIdx.x];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = v;
    }
}

// single block: exclusive scan of up to 1024*1024/kScanBlock block sums, serially chunked
__global__ void k_scan_of_block_sums(unsigned int* __restrict__ block_sums, int n_blocks) {
    __shared__ unsigned int sh[kScanBlock];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += kScanBlock) {
        const int i = base + threadIdx.x;
        const unsigned int v = (i < n_blocks) ? block_sums[i] : 0u;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < kScanBlock; o <<= 1) {       // Hillis-Steele inclusive scan
            unsigned int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0u;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n_blocks) block_sums[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += sh[kScanBlock - 1];
        __syncthreads();
    }
}

// offsets[i] = exclusive prefix of counts; cursor[i] = offsets[i] (scatter cursor); mask[pixel] = count > 0
__global__ void k_scan_finish(const unsigned int* __restrict__ counts, int n_keys, const unsigned int* __restrict__ block_sums,
                              unsigned int* __restrict__ cursor) {
    __shared__ unsigned int sh[kScanBlock];
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    const unsigned int v = (i < n_keys) ? counts[i] : 0u;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < kScanBlock; o <<= 1) {
        unsigned int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0u;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n_keys) cursor[i] = block_sums[blockIdx.x] + sh[threadIdx.x] - v;
}

// Scatter events into pixel-sorted order.  ev_xy packs (x | y << 16); perm keeps the original index
// (needed by the debug index tap and by k_rank_sort_segments).  Order inside one pixel is arbitrary here.
__global__ void k_scatter_events(const int16_t* __restrict__ xs, const int16_t* __restrict__ ys, const double* __restrict__ ts,
                                 int64_t n, int H, int W, int tiles_x, unsigned int* __restrict__ cursor,
                                 uint32_t* __restrict__ ev_xy, double* __restrict__ ev_t, uint32_t* __restrict__ perm) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int x = xs[e], y = ys[e];
        if (x < 0 || x >= W || y < 0 || y >= H) continue;
        const unsigned int pos = atomicAdd(&cursor[sort_key(x, y, W, tiles_x)], 1u);
        ev_xy[pos] = (uint32_t)x | ((uint32_t)y << 16);
        ev_t[pos] = ts[e];
        perm[pos] = (uint32_t)e;
    }
}

// Orders the events of every pixel by their original index (= by time: the loaders deliver ts ascending,
// reference src/dataloaders/dsec_loader.py:285-349), turning the unstable scatter above into a stable sort.  Consecutive
// sorted events then differ little in (pixel, t), so their warped destinations coincide often and the event kernels can
// merge them in registers before touching memory (k_events9.cuh).  Rank sort: one thread per event counts the events of
// its own pixel segment with a smaller original index.  Segments longer than kMaxRankSeg (hot pixels) keep the scatter
// order - the order only affects how many reductions are merged, never the result.
constexpr unsigned int kMaxRankSeg = 4096;

__global__ void k_rank_sort_segments(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, const uint32_t* __restrict__ perm,
                                     int64_t n, int W, int tiles_x, const unsigned int* __restrict__ counts,
                                     const unsigned int* __restrict__ cursor_end, double* __restrict__ ev_t_out,
                                     uint32_t* __restrict__ perm_out) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t xy = ev_xy[e];
        const int key = sort_key(xy & 0xffffu, xy >> 16, W, tiles_x);
        const unsigned int cnt = counts[key], end = cursor_end[key], start = end - cnt;
        const uint32_t p = perm[e];
        unsigned int rank = (unsigned int)e - start;
        if (cnt > 1u && cnt <= kMaxRankSeg) {
            rank = 0;
            for (unsigned int j = start; j < end; ++j) rank += perm[j] < p ? 1u : 0u;
        }
        ev_t_out[start + rank] = ev_t[e];
        perm_out[start + rank] = p;
    }
}

// mask[y*W+x] = 1 where the pixel holds at least one event (reference src/utils/theta_utils.py:66-71)
__global__ void k_event_mask(const unsigned int* __restrict__ counts, int H, int W, int tiles_x, uint8_t* __restrict__ mask) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= H * W) return;
    const int y = p / W, x = p % W;
    mask[p] = counts[sort_key(x, y, W, tiles_x)] > 0u ? 1 : 0;
}

}  // namespace eincm
