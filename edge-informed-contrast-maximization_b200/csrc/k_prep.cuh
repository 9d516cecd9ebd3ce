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

// ---- layout of the sorted stream ---------------------------------------------------------------------------------
// Every 16x16 source tile owns one contiguous segment of the sorted stream, padded with kNoEvent sentinels to a multiple of
// kStreamAlign events (so that a thread's vector load of 4 events never straddles two tiles), and is cut into chunks of
// <= kChunkEvents events - the work items of the tile-privatised event kernels (k_events_tile.cuh).
constexpr int kKeysPerTile = kSortTile * kSortTile;     // 256 sort keys (pixels) per tile
constexpr unsigned int kStreamAlign = 4;
// Several CTA passes of 1024 events (4 per thread) share one window set-up, one zeroing and one flush: most tiles of a DSEC window hold
// more than 1024 events, and the per-chunk overhead was half of the splat's instructions.  A window cell is a uint32 of
// 2^kFixShift-scaled votes: a chunk holds at most 2^(32 - kFixShift) - 4 events, so that the cell cannot overflow even if every event of
// the chunk put a full centre tap on it.
constexpr int kFixShift = 21;     // (20 would allow 4092-event chunks: measured 2 us of 128 per DSEC evaluation, for half the vote resolution)
constexpr unsigned int kChunkEvents = (1u << (32 - kFixShift)) - 4u;

struct Chunk { uint32_t start, count, origin, pad; };   // count: multiple of kStreamAlign, <= kChunkEvents; origin: tile corner x | y << 16

// tile_cnt[t] = number of events of tile t
__global__ void __launch_bounds__(kKeysPerTile)
k_tile_counts(const unsigned int* __restrict__ counts, unsigned int* __restrict__ tile_cnt) {
    __shared__ unsigned int sh[kKeysPerTile / 32];
    unsigned int v = counts[blockIdx.x * kKeysPerTile + threadIdx.x];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int s = 0;
        for (int k = 0; k < kKeysPerTile / 32; ++k) s += sh[k];
        tile_cnt[blockIdx.x] = s;
    }
}

// single CTA of 1024 threads: exclusive scans over the tiles of (padded event count, chunk count), 1024 tiles per round.
// totals[0] = padded stream length, totals[1] = number of chunks.
__global__ void __launch_bounds__(1024)
k_tile_layout(const unsigned int* __restrict__ tile_cnt, int n_tiles, unsigned int* __restrict__ tile_start,
              unsigned int* __restrict__ chunk_first, unsigned int* __restrict__ totals) {
    __shared__ unsigned int sh_a[32], sh_b[32];
    __shared__ unsigned int carry_a, carry_b;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry_a = 0; carry_b = 0; }
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int t = base + threadIdx.x;
        const unsigned int cnt = t < n_tiles ? tile_cnt[t] : 0u;
        const unsigned int a = (cnt + kStreamAlign - 1) / kStreamAlign * kStreamAlign;
        const unsigned int b = (cnt + kChunkEvents - 1) / kChunkEvents;
        unsigned int ia = a, ib = b;                         // inclusive warp scans
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int ua = __shfl_up_sync(0xffffffffu, ia, o), ub = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) { ia += ua; ib += ub; }
        }
        if (lane == 31) { sh_a[wid] = ia; sh_b[wid] = ib; }
        __syncthreads();
        if (wid == 0) {
            unsigned int wa = sh_a[lane], wb = sh_b[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int ua = __shfl_up_sync(0xffffffffu, wa, o), ub = __shfl_up_sync(0xffffffffu, wb, o);
                if (lane >= o) { wa += ua; wb += ub; }
            }
            sh_a[lane] = wa; sh_b[lane] = wb;                // inclusive over warps
        }
        __syncthreads();
        const unsigned int off_a = carry_a + (wid ? sh_a[wid - 1] : 0u) + ia - a;
        const unsigned int off_b = carry_b + (wid ? sh_b[wid - 1] : 0u) + ib - b;
        if (t < n_tiles) { tile_start[t] = off_a; chunk_first[t] = off_b; }
        __syncthreads();
        if (threadIdx.x == 0) { carry_a += sh_a[31]; carry_b += sh_b[31]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { tile_start[n_tiles] = carry_a; chunk_first[n_tiles] = carry_b; totals[0] = carry_a; totals[1] = carry_b; }
}

// one CTA per tile: cursor[key] = tile_start + exclusive prefix of the tile's key counts; chunk records of the tile
__global__ void __launch_bounds__(kKeysPerTile)
k_tile_finish(const unsigned int* __restrict__ counts, const unsigned int* __restrict__ tile_cnt, const unsigned int* __restrict__ tile_start,
              const unsigned int* __restrict__ chunk_first, int tiles_x, unsigned int* __restrict__ cursor, Chunk* __restrict__ chunks) {
    __shared__ unsigned int sh[kKeysPerTile / 32];
    const int t = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned int v = counts[t * kKeysPerTile + threadIdx.x];
    unsigned int iv = v;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int u = __shfl_up_sync(0xffffffffu, iv, o);
        if (lane >= o) iv += u;
    }
    if (lane == 31) sh[wid] = iv;
    __syncthreads();
    unsigned int before = 0;
    for (int k = 0; k < wid; ++k) before += sh[k];
    const unsigned int start = tile_start[t];
    cursor[t * kKeysPerTile + threadIdx.x] = start + before + iv - v;
    const unsigned int cnt = tile_cnt[t];
    const unsigned int padded = (cnt + kStreamAlign - 1) / kStreamAlign * kStreamAlign;
    const unsigned int nch = (cnt + kChunkEvents - 1) / kChunkEvents, first = chunk_first[t];
    for (unsigned int j = threadIdx.x; j < nch; j += kKeysPerTile) {
        Chunk c;
        c.start = start + j * kChunkEvents;
        c.count = min(kChunkEvents, padded - j * kChunkEvents);
        c.origin = (uint32_t)((t % tiles_x) * kSortTile) | ((uint32_t)((t / tiles_x) * kSortTile) << 16);
        c.pad = 0u;
        chunks[first + j] = c;
    }
}

// Chunk table in descending order of event count (ties by table position: deterministic).  The event kernels run one CTA per chunk and
// the hardware hands CTAs out in index order, so this is longest-processing-time-first list scheduling: the tail of the grid is
// filled with the small chunks instead of whatever tiles happen to be last in the image.  One thread per chunk counts the chunks
// that precede it (n^2 / 2 comparisons on L1-resident data; once per window; the host skips tables beyond kMaxOrderedChunks).
constexpr int kMaxOrderedChunks = 16384;

__global__ void __launch_bounds__(256)
k_chunk_order(const Chunk* __restrict__ in, const unsigned int* __restrict__ n_chunks_dev, Chunk* __restrict__ out) {
    const int n = (int)__ldg(n_chunks_dev);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Chunk me = in[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
        const unsigned int cj = __ldg(&in[j].count);
        rank += (cj > me.count || (cj == me.count && j < i)) ? 1 : 0;
    }
    out[rank] = me;
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
        if (xy == 0xffffffffu) continue;                    // padding sentinel (n = capacity of the padded stream)
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
