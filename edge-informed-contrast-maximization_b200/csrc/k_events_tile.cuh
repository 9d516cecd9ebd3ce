// Tile-privatised event-space kernels (default path).
//
// The staged event stream is sorted by source pixel, tile-major over 16x16 source tiles, and by time inside a pixel
// (k_prep.cuh); every tile's segment is padded to a multiple of kEvK events and cut into chunks of <= kChunk events.
// One CTA owns one chunk.  Because all events of a chunk start inside one 16x16 tile, their warped 3x3 patches
// (reference src/utils/event_utils.py:41-59) fall into a small rectangle of the destination image per reference time:
// the CTA bounds that rectangle (interval arithmetic on the tile's theta range and the chunk's t range), keeps it as a window in shared memory, votes into it
// with native 32-bit shared-memory integer atomics (ATOMS.ADD), and finally adds the non-zero window cells to the
// global image with one 64-bit integer reduction each.  Votes are fixed point, 2^kFixShift * (2 pi v):
//   * shared-memory float atomics are CAS loops on sm_100a, integer adds are native (see profiles/microbench);
//   * integer sums do not depend on the order of the votes, so the image of warped events - and therefore the
//     objective - is bit-reproducible from run to run, which matters to BFGS with gtol = 1e-7 (main.yaml:34).
// A vote is quantised to 2^-21 of the centre-tap value (4.8e-7, about 4 float32 ulps of the largest tap); coordinates,
// the warp and rint() stay float64, so pixel indices are bit-exact.  A rectangle larger than kWinCap cells (flows beyond
// ~45 px per window) is processed in row SLICES, one more pass over the chunk per RB slices; events outside a rectangle
// cropped at kMaxSlices slices / kWinMaxW columns, or exactly on a rounding boundary between slices, fall back to per-tap
// global reductions.  The reference's index rule (negative indices wrap, out of range drops: SURVEY.md A.4) is applied
// once per window cell at flush time.
//
// Backward: the same rectangles (recorded by the forward pass) are filled with d loss / d IWE and the nine taps of every
// event are read from shared memory instead of global memory; of a sliced rectangle the backward pass keeps the first
// slice and gathers the rest from the global image (it only reads: measured cheaper than further passes).
#pragma once
#include <climits>
#include <type_traits>

#include "common.cuh"
#include "k_events.cuh"
#include "k_prep.cuh"
#include "k_theta.cuh"

namespace eincm {

constexpr int kEvK = 4;                    // events per thread (one 128-bit load of packed coordinates, two of timestamps)
constexpr uint32_t kNoEvent = 0xffffffffu; // padding sentinel of the sorted stream
constexpr int kMaxRB = 4;                  // reference times processed per pass over a chunk
// reference times per pass for R reference times: the passes a kMaxRB-wide kernel would need, evenly filled (R = 5: two passes of 3 instead of
// 4 + 1 - the same number of passes with 48 KB instead of 64 KB of windows per CTA, i.e. four resident CTAs per SM instead of three)
inline int refs_per_pass(int R) {
    const int passes = (R + kMaxRB - 1) / kMaxRB;
    return passes > 0 ? (R + passes - 1) / passes : 1;
}

// Warp of one event to one reference time on the default path.  Same float64 arithmetic, in the same order, as warp_event
// (event_warpers.py:34-35: x' = x - (theta * dt) * 1.0), but rint() and the int conversion use the 2^52 magic constant
// (two DADDs instead of F2I + I2F on the slow conversion pipe): for |x'| < 2^31, (x' + M) - M == rint(x') under
// round-half-to-even and the low word of (x' + M) is that integer.  `ok` is false for non-finite / absurdly far warps
// (every tap is out of range under either index rule: dropped).
struct Hit { int rx, ry; float fx, fy; bool ok; };

__device__ __forceinline__ Hit warp_hit(double xd, double yd, double thx, double thy, double dt) {
    constexpr double kMagic = 6755399441055744.0;   // 1.5 * 2^52
    const double xw = __dsub_rn(xd, __dmul_rn(thx, dt));
    const double yw = __dsub_rn(yd, __dmul_rn(thy, dt));
    const double sx = __dadd_rn(xw, kMagic), sy = __dadd_rn(yw, kMagic);
    Hit h;
    h.ok = (fabs(xw) < 1.0e9) && (fabs(yw) < 1.0e9);      // false for NaN
    h.rx = __double2loint(sx);
    h.ry = __double2loint(sy);
    h.fx = (float)__dsub_rn(xw, __dsub_rn(sx, kMagic));
    h.fy = (float)__dsub_rn(yw, __dsub_rn(sy, kMagic));
    return h;
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

__device__ __forceinline__ void red_G(double* __restrict__ G, int W, uint32_t xy, float sx, float sy) {
    double* g = G + ((int64_t)(xy >> 16) * W + (xy & 0xffffu)) * 2;
    atomicAdd(g, (double)sx);
    atomicAdd(g + 1, (double)sy);
}


constexpr int kSubChunk = kEvK * 256;       // events a CTA holds in registers at a time (kEvK per thread): a chunk is processed in sub-chunks
static_assert(kChunkEvents % kEvK == 0 && kStreamAlign == kEvK, "chunks are made of groups of kEvK events");
static_assert((unsigned long long)kChunkEvents << kFixShift < (1ull << 32), "a uint32 window cell holds the votes of a whole chunk");
constexpr int kWinCap = 4096;         // window cells per reference time (16 KB of uint32 / float)
constexpr double kFixToIwe = kInv2Pi / (double)(1u << kFixShift);        // fixed-point sum -> image value
// float -> fixed point by mantissa alignment: for 0 <= p <= 1, the float 6 + p lies in [4, 8) where one ulp is 2^-21, so
// bits(fma(ex, ey, 6.0f)) - bits(6.0f) == round(2^21 * ex * ey) (round to nearest even), no scaling multiply
static_assert(kFixShift == 21, "the rounding constant below is chosen for 2^-21");
constexpr float kRoundMagic = 6.0f;
constexpr int kRoundMagicBits = 0x40C00000;

// exp(-0.5 (d - f)^2) for d = -1, 0, 1:  u = s f, q_d = s d - u, value = 2^(-q_d^2) with s = sqrt(0.5 log2 e)
__device__ __forceinline__ void axis_exp3(float f, float e[3]) {
    constexpr float s = 0.84932180028801907f;
    const float u = s * f;
    const float q0 = -s - u, q2 = s - u;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[0]) : "f"(-q0 * q0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[1]) : "f"(-u * u));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[2]) : "f"(-q2 * q2));
}

struct TapsFix { int n[9]; };                            // index (j+1)*3 + (i+1): column offset i, row offset j

// nine fixed-point tap values of one warped event: round(2^21 * exp(-0.5 ((i - fx)^2 + (j - fy)^2)))
__device__ __forceinline__ TapsFix taps_fix(float fx, float fy) {
    float ex[3], ey[3];
    axis_exp3(fx, ex);
    axis_exp3(fy, ey);
    TapsFix t;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) t.n[j * 3 + i] = __float_as_int(fmaf(ex[i], ey[j], kRoundMagic)) - kRoundMagicBits;
    return t;
}

// Thread -> event group (kEvK consecutive events of the sorted stream) of one SUB-CHUNK.  A chunk of up to kChunkEvents events is
// processed kSubChunk events at a time (its groups split evenly over the sub-chunks).  The backward pass gives consecutive groups to
// consecutive lanes (neighbouring lanes hold the same or neighbouring source pixels: its per-pixel segmented sums run over lanes, and
// equal shared-memory addresses are broadcasts for loads).  The splat spreads the lanes of a warp over the sub-chunk - with nw = warps
// needed for its groups, lane l of warp w < nw takes group nw * l + w: consecutive events land on the same or neighbouring destination
// cells, and equal addresses inside one shared-memory ATOMIC instruction are serialised.
__device__ __forceinline__ int n_subs(const Chunk ch) { return (int)((ch.count + kSubChunk - 1) / kSubChunk); }

template <bool SPREAD>
__device__ __forceinline__ void load_sub_events(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, const Chunk ch, int sub,
                                                EventGroup& ev) {
    const uint32_t G = ch.count >> 2, ns = (uint32_t)n_subs(ch), per = (G + ns - 1u) / max(ns, 1u);     // groups per sub-chunk (<= 256)
    const uint32_t g0 = min(G, (uint32_t)sub * per), ng = min(G, g0 + per) - g0;
    uint32_t g = threadIdx.x;
    if (SPREAD) {
        const uint32_t nw = (ng + 31u) >> 5, w = threadIdx.x >> 5;                  // 32 groups (128 events) per full warp
        g = w < nw ? nw * (threadIdx.x & 31u) + w : 0xffffu;
    }
    if (g < ng) {
        load_group(ev_xy, ev_t, (int64_t)(ch.start >> 2) + g0 + g, ev);
    } else {
#pragma unroll
        for (int k = 0; k < kEvK; ++k) { ev.xy[k] = kNoEvent; ev.t[k] = 0.0; }
    }
}

// Small chunks (sparse windows: MVSEC has ~90 events per source tile) give every thread ONE event instead of kEvK: the serial work of a thread
// between two barriers of the chunk - the critical path of a CTA in which seven of eight warps would otherwise wait at the barrier for one
// (ncu, profiles/r2_ncu_metrics_mvsec_batch.txt: 9.9 barrier stalls per issue in k_backward_tile_b) - shrinks by kEvK.  Slot 0 of the
// group holds the event, the loops over the slots stop after it.  Same lane order as load_sub_events.
#ifndef EINCM_SINGLE_MAX
#define EINCM_SINGLE_MAX 256
#endif
constexpr uint32_t kSingleMax = EINCM_SINGLE_MAX;   // 0: never (A/B builds)
__device__ __forceinline__ bool single_mode(const Chunk ch) { return ch.count <= kSingleMax; }

template <bool SPREAD>
__device__ __forceinline__ void load_single_events(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, const Chunk ch, EventGroup& ev) {
    uint32_t e = threadIdx.x;
    if (SPREAD) {
        const uint32_t nw = (ch.count + 31u) >> 5, w = threadIdx.x >> 5;
        e = w < nw ? nw * (threadIdx.x & 31u) + w : 0xffffffffu;
    }
#pragma unroll
    for (int k = 0; k < kEvK; ++k) { ev.xy[k] = kNoEvent; ev.t[k] = 0.0; }
    if (e < ch.count) {
        ev.xy[0] = __ldg(ev_xy + (int64_t)ch.start + e);
        ev.t[0] = __ldg(ev_t + (int64_t)ch.start + e);
    }
}

// Window of one reference time inside the destination image: origin (ox, oy), pw x ph cells, row pitch pw.  A warped event
// votes into the window when its rounded centre lies in [ox + 1, ox + pw - 2] x [oy + 1, oy + ph - 2]; the kernels test that on
// the float64 warped coordinate itself, |x' - cx| < hx (strict: a coordinate exactly on the rounding boundary takes the
// fallback), which also rejects NaN / infinite / absurdly far warps in the same two comparisons.
struct Window {
    int ox, oy, pw, ph; double cx, hx, cy, hy; uint32_t inv_pw, interior;   // inv_pw = ceil(2^32 / pw): exact i / pw for i < 2^16
    double tr; int img, pad;                                                // reference time and image the window belongs to
};

__device__ __forceinline__ Window make_window(int ox, int oy, int pw, int ph) {
    Window w;
    w.ox = ox; w.oy = oy; w.pw = pw; w.ph = ph;
    w.cx = (double)ox + 0.5 * (double)(pw - 1); w.hx = 0.5 * (double)(pw - 2);
    w.cy = (double)oy + 0.5 * (double)(ph - 1); w.hy = 0.5 * (double)(ph - 2);
    if (pw < 3 || ph < 3) { w.hx = -1.0; w.hy = -1.0; }
    w.inv_pw = pw > 0 ? 0xffffffffu / (uint32_t)pw + 1u : 0u;
    w.interior = 0u;
    w.tr = 0.0; w.img = 0; w.pad = 0;
    return w;
}

// every cell of the window lies inside the H x W image: the index rule of the reference (wrap / drop) has nothing to do
__device__ __forceinline__ void set_interior(Window& w, int H, int W) {
    w.interior = (w.ox >= 0 && w.oy >= 0 && w.ox + w.pw <= W && w.oy + w.ph <= H) ? 1u : 0u;
}

// Destination rectangle of one (chunk, reference time) as (ox, oy, pw, ph): holds the centres [mnx, mxx] x [mny, mxy] plus a halo
// of one cell.  A rectangle larger than a shared-memory window (kWinCap cells: large flows) is processed in SLICES of complete
// rows, one pass over the chunk's events per slice, each slice being an ordinary window (centre rows partitioned, halo rows
// shared); only rectangles wider than kWinMaxW cells or taller than kMaxSlices slices are cropped - the events outside take
// the per-tap global fallback.
constexpr int kWinMaxW = kWinCap / 3;   // a window needs at least three rows
constexpr int kMaxSlices = 16;
constexpr int kMissThreadsPerPass = 0;    // more threads (of 256) than this with left-over events justify one more pass over the chunk

__device__ __forceinline__ int slice_rows(int pw) { return kWinCap / pw - 2; }       // centre rows per slice of a pw-wide rectangle

__device__ __forceinline__ int4 bound_rect(int mnx, int mny, int mxx, int mxy) {
    long long bw = (long long)mxx - mnx + 3, bh = (long long)mxy - mny + 3;
    if (bw > kWinMaxW) bw = kWinMaxW;
    if (bw * bh > kWinCap) {
        const long long max_h = (long long)slice_rows((int)bw) * kMaxSlices + 2;
        if (bh > max_h) bh = max_h;
    }
    return make_int4(mnx - 1, mny - 1, (int)bw, (int)bh);
}

__device__ __forceinline__ int n_slices(const int4 rect) {
    if (rect.z < 3 || rect.w < 3) return 0;
    if (rect.z * rect.w <= kWinCap) return 1;
    const int hs = slice_rows(rect.z);
    return (rect.w - 2 + hs - 1) / hs;
}

// window of slice s of a rectangle (an empty window when the rectangle has fewer slices)
__device__ __forceinline__ Window slice_window(const int4 rect, int s) {
    if (s >= n_slices(rect)) return make_window(0, 0, 0, 0);
    if (rect.z * rect.w <= kWinCap) return make_window(rect.x, rect.y, rect.z, rect.w);
    const int hs = slice_rows(rect.z);
    const int first = rect.y + 1 + s * hs;                                  // first centre row of the slice
    const int n = min(hs, rect.y + rect.w - 1 - first);                     // centre rows: up to rect.y + ph - 2
    return make_window(rect.x, first - 1, rect.z, n + 2);
}

// A pass of the event kernels holds RB windows.  The windows of a chunk are its (reference time, slice) PAIRS: pair j < R is the first
// slice of reference time j (pairs R .. Rpad - 1 are empty, Rpad = R rounded up to a multiple of RB), the further slices of sliced
// rectangles follow in reference-major order from pair Rpad: off[r] = index of the pair of slice 1 of reference r, off[R] = number
// of pairs (= Rpad unless a rectangle is sliced).  Window of pair j (an empty window where there is none).
__device__ __forceinline__ Window pair_window(const int4* rects, const int* off, int R, int Rpad, int j, const RefTimes& tref, int H, int W) {
    Window wn = make_window(0, 0, 0, 0);
    if (j < R) {
        wn = slice_window(rects[j], 0);
        wn.tr = tref.t[j];
        wn.img = j;
    } else if (j >= Rpad && j < off[R]) {
        int r = 0;
        while (off[r + 1] <= j) ++r;
        wn = slice_window(rects[r], j - off[r] + 1);
        wn.tr = tref.t[r];
        wn.img = r;
    }
    set_interior(wn, H, W);
    return wn;
}

// slice counts of the lanes' rectangles -> off[] (warp-wide, lanes >= R pass ns = 0)
__device__ __forceinline__ void pair_offsets(int ns, int R, int Rpad, int* off) {
    const int lane = threadIdx.x & 31;
    int incl = max(ns - 1, 0);
#pragma unroll
    for (int o = 1; o < EINCM_MAX_REFS; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane < R) off[lane + 1] = Rpad + incl;
    if (lane == 0) off[0] = Rpad;
}

// true when the warped coordinate votes into one of the slices of the rectangle (the hit test of the hot loops, replayed
// out of line: a coordinate exactly on the rounding boundary between two slices belongs to neither and takes the fallback)
__device__ __noinline__ bool hits_rect(const int4 rect, double xw, double yw, int max_slices) {
    const int ns = min(n_slices(rect), max_slices);
    for (int s = 0; s < ns; ++s) {
        const Window w = slice_window(rect, s);
        if ((fabs(xw - w.cx) < w.hx) & (fabs(yw - w.cy) < w.hy)) return true;
    }
    return false;
}

// float <-> int32 with the same ordering (for REDUX.MIN / REDUX.MAX); NaN maps beyond +-inf and survives the round trip
__device__ __forceinline__ int ordered_int(float f) { const int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// Conservative destination rectangle of one chunk and one reference time: every event of the chunk starts inside the 16x16 source tile at
// (x0, y0), has theta inside [thx_lo, thx_hi] x [thy_lo, thy_hi] (range over the tile) and t inside [t_lo, t_hi] (range over
// the chunk, k_chunk_trange), so x' = x - theta_x (t - t_ref) lies inside an interval known without touching the events.
// float32 interval arithmetic with explicit slack; rint(x') of every event is inside [floor(lo + 0.49), ceil(hi - 0.49)].
__device__ __forceinline__ int4 chunk_rect(int x0, int y0, int x1, int y1, float thx_lo, float thx_hi, float thy_lo, float thy_hi,
                                               float t_lo, float t_hi, double t_ref) {
    const float trf = (float)t_ref;
    const float es = 1.0e-6f * (1.f + fabsf(trf) + fmaxf(fabsf(t_lo), fabsf(t_hi)));
    const float d_lo = (t_lo - trf) - es, d_hi = (t_hi - trf) + es;
    const float px_lo = fminf(fminf(thx_lo * d_lo, thx_lo * d_hi), fminf(thx_hi * d_lo, thx_hi * d_hi));
    const float px_hi = fmaxf(fmaxf(thx_lo * d_lo, thx_lo * d_hi), fmaxf(thx_hi * d_lo, thx_hi * d_hi));
    const float py_lo = fminf(fminf(thy_lo * d_lo, thy_lo * d_hi), fminf(thy_hi * d_lo, thy_hi * d_hi));
    const float py_hi = fmaxf(fmaxf(thy_lo * d_lo, thy_lo * d_hi), fmaxf(thy_hi * d_lo, thy_hi * d_hi));
    const float lox = (float)x0 - px_hi, hix = (float)x1 - px_lo, loy = (float)y0 - py_hi, hiy = (float)y1 - py_lo;
    const float mx = 1.0e-5f * (fabsf(lox) + fabsf(hix)) + 1.0e-3f, my = 1.0e-5f * (fabsf(loy) + fabsf(hiy)) + 1.0e-3f;
    // the comparison is false for NaN (theta or t not finite): no window, every event of the chunk takes the fallback
    const bool finite = (lox > -1.0e9f) && (hix < 1.0e9f) && (loy > -1.0e9f) && (hiy < 1.0e9f) &&
                        (fabsf(thx_lo) + fabsf(thx_hi) + fabsf(thy_lo) + fabsf(thy_hi) < 3.0e38f) && (d_hi - d_lo < 3.0e38f);
    if (!finite) return make_int4(0, 0, 0, 0);
    return bound_rect(__float2int_rd(lox + 0.49f - mx), __float2int_rd(loy + 0.49f - my),
                        __float2int_ru(hix - 0.49f + mx), __float2int_ru(hiy - 0.49f + my));
}

// per window: t range of every chunk (float32, rounded outwards; +inf / -inf for a chunk without events)
__global__ void __launch_bounds__(256)
k_chunk_trange(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, const Chunk* __restrict__ chunks,
               const unsigned int* __restrict__ n_chunks_dev, float2* __restrict__ chunk_tr) {
    __shared__ int sh[8][2];
    const int n_chunks = (int)__ldg(n_chunks_dev);
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const Chunk ch = chunks[c];
        float lo = INFINITY, hi = -INFINITY;
        bool nan = false;                       // fminf / fmaxf drop NaN operands: a NaN timestamp must poison the range instead
        for (int sub = 0; sub < n_subs(ch); ++sub) {
            EventGroup ev;
            load_sub_events<false>(ev_xy, ev_t, ch, sub, ev);
#pragma unroll
            for (int k = 0; k < kEvK; ++k)
                if (ev.xy[k] != kNoEvent) {
                    nan |= ev.t[k] != ev.t[k];
                    lo = fminf(lo, __double2float_rd(ev.t[k]));
                    hi = fmaxf(hi, __double2float_ru(ev.t[k]));
                }
        }
        const float qnan = __int_as_float(0x7fc00000);
        const int rl = __reduce_min_sync(0xffffffffu, ordered_int(nan ? -qnan : lo));     // -NaN orders below -inf
        const int rh = __reduce_max_sync(0xffffffffu, ordered_int(nan ? qnan : hi));      // +NaN orders above +inf
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5][0] = rl; sh[threadIdx.x >> 5][1] = rh; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int a = sh[0][0], b = sh[0][1];
            for (int w8 = 1; w8 < 8; ++w8) { a = min(a, sh[w8][0]); b = max(b, sh[w8][1]); }
            chunk_tr[c] = make_float2(ordered_float(a), ordered_float(b));
        }
    }
}

// Adds nine tap values to the 3x3 cells around the shared-memory cell `mid` (rows `pitch4` bytes apart): nine
// `red.shared.add.u32` (ATOMS.ADD without return value) with immediate column offsets.
__device__ __forceinline__ void emit9(uint32_t mid, uint32_t pitch4, const int (&a)[9]) {
    const uint32_t up = mid - pitch4, dn = mid + pitch4;
    asm volatile(
        "red.shared.add.u32 [%0 + -4], %3;\n\tred.shared.add.u32 [%0], %4;\n\tred.shared.add.u32 [%0 + 4], %5;\n\t"
        "red.shared.add.u32 [%1 + -4], %6;\n\tred.shared.add.u32 [%1], %7;\n\tred.shared.add.u32 [%1 + 4], %8;\n\t"
        "red.shared.add.u32 [%2 + -4], %9;\n\tred.shared.add.u32 [%2], %10;\n\tred.shared.add.u32 [%2 + 4], %11;"
        ::"r"(up), "r"(mid), "r"(dn), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8])
        : "memory");
}

// theta of the 16x16 source tile at `origin` -> shared memory (zero flow when T.theta is null); returns the thread's value
__device__ __forceinline__ double2 tile_theta(const ThetaSrc& T, uint32_t origin, int H, int W, double2* __restrict__ th_s) {
    const int p = threadIdx.x;                       // 256 threads <-> 256 pixels
    const int x = (int)(origin & 0xffffu) + (p & 15), y = (int)(origin >> 16) + (p >> 4);
    double2 v = make_double2(0.0, 0.0);
    if (T.theta != nullptr && x < W && y < H) v = theta_at(T, x, y);
    th_s[p] = v;
    return v;
}

// range of theta over the tile: per-warp REDUX of the order-preserving integer images of (float) theta -> sbox[warp][0..3]
__device__ __forceinline__ void tile_theta_range(const double2 v, int (*sbox)[4]) {
    const int ax = __reduce_min_sync(0xffffffffu, ordered_int(__double2float_rd(v.x)));
    const int bx = __reduce_max_sync(0xffffffffu, ordered_int(__double2float_ru(v.x)));
    const int ay = __reduce_min_sync(0xffffffffu, ordered_int(__double2float_rd(v.y)));
    const int by = __reduce_max_sync(0xffffffffu, ordered_int(__double2float_ru(v.y)));
    if ((threadIdx.x & 31) == 0) { int* s = sbox[threadIdx.x >> 5]; s[0] = ax; s[1] = bx; s[2] = ay; s[3] = by; }
}

// shared-memory address that the compiler cannot rematerialise (it would rebuild it from %cluster_ctarank at every use)
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("mov.u32 %0, %0;" : "+r"(a));
    return a;
}

__device__ __forceinline__ double2 lds_theta(uint32_t th_base, uint32_t xy) {
    double2 v;                                              // ((y & 15) * 16 + (x & 15)) * 16 bytes
    const uint32_t a = th_base + ((((xy >> 12) & 0xf0u) | (xy & 0xfu)) << 4);
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}

// Destination images of the flush.  Single GPU: the plan's own fixed-point image.  Event split with peer access: the images
// of ALL ranks of the split (own + peers' device pointers opened through CUDA IPC) - every rank adds its non-zero window
// cells to every rank's image with 64-bit integer reductions over NVLink, so the "all-reduce of the partial images" is fused
// into the splat: integer sums are order-independent, every rank ends up with the identical complete image.
constexpr int kMaxPeers = 8;
struct FixDst { unsigned long long* p[kMaxPeers]; int n; };

// float64 warp of one event, shared by the hot loops and the cold paths (event_warpers.py:34-35: x' = x - (theta * dt) * 1.0).
// rint() and the int conversion use the 2^52 magic constant (two DADDs instead of F2I + I2F on the slow conversion pipe): for
// |x'| < 2^31, (x' + M) - M == rint(x') under round-half-to-even and the low word of (x' + M) is that integer.
struct Hit2 { double xw, yw; int rx, ry; float fx, fy; };

__device__ __forceinline__ Hit2 warp_hit2(uint32_t xy, double2 th, double dt) {
    constexpr double kMagic = 6755399441055744.0;   // 1.5 * 2^52
    Hit2 h;
    h.xw = __dsub_rn((double)(xy & 0xffffu), __dmul_rn(th.x, dt));
    h.yw = __dsub_rn((double)(xy >> 16), __dmul_rn(th.y, dt));
    const double sx = __dadd_rn(h.xw, kMagic), sy = __dadd_rn(h.yw, kMagic);
    h.rx = __double2loint(sx);
    h.ry = __double2loint(sy);
    h.fx = (float)__dsub_rn(h.xw, __dsub_rn(sx, kMagic));
    h.fy = (float)__dsub_rn(h.yw, __dsub_rn(sy, kMagic));
    return h;
}

// cold path of the splat: an event whose patch leaves its window (or whose warp is not finite) adds its taps to the global
// images one by one, with the reference's index rule; non-finite / absurdly far warps are dropped (every tap is out of range
// under either index rule).  Out of line and self-contained: the hot loop keeps no state alive for it.
template <bool WRAP>
__device__ __noinline__ void splat_fallback(const FixDst& dst, int64_t img_off, uint32_t xy, double2 th, double dt, int H, int W) {
    const Hit2 h = warp_hit2(xy, th, dt);
    if (!((fabs(h.xw) < 1.0e9) && (fabs(h.yw) < 1.0e9))) return;
    // the whole patch is out of range under the index rule (line-search trial steps of BFGS reach flows of thousands of pixels,
    // where that is every event): nothing to add, no tap values needed
    if (h.rx + 1 < (WRAP ? -W : 0) || h.rx - 1 >= W || h.ry + 1 < (WRAP ? -H : 0) || h.ry - 1 >= H) return;
    const TapsFix t = taps_fix(h.fx, h.fy);
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) {
            int rr = h.ry + j - 1, cc = h.rx + i - 1;
            if (drop_index<WRAP>(rr, cc, H, W)) {
                const int64_t off = img_off + (int64_t)rr * W + cc;
#pragma unroll 1
                for (int q = 0; q < dst.n; ++q) atomicAdd(dst.p[q] + off, (unsigned long long)t.n[j * 3 + i]);
            }
        }
}

// d loss / d x', d loss / d y' of one warped event from the nine cotangent cells d[] (already scaled by 1 / 2 pi):
// separable evaluation  wx_i = exp(-0.5 (i - fx)^2), s_j = sum_i D_ij wx_i, sx_j = sum_i D_ij wx_i (i - fx)
//   dL/dx' = sum_j wy_j sx_j,   dL/dy' = sum_j wy_j (j - fy) s_j
__device__ __forceinline__ void tap_gradient(const float (&d)[9], float fx, float fy, float& gx, float& gy) {
    float wx[3], wy[3];
    axis_exp3(fx, wx);
    axis_exp3(fy, wy);
    const float ux0 = wx[0] * (-1.f - fx), ux1 = -fx * wx[1], ux2 = wx[2] * (1.f - fx);
    float sj[3], sxj[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        sj[j] = fmaf(d[j * 3 + 0], wx[0], fmaf(d[j * 3 + 2], wx[2], d[j * 3 + 1] * wx[1]));
        sxj[j] = fmaf(d[j * 3 + 0], ux0, fmaf(d[j * 3 + 2], ux2, d[j * 3 + 1] * ux1));
    }
    gx = fmaf(wy[0], sxj[0], fmaf(wy[2], sxj[2], wy[1] * sxj[1]));
    gy = fmaf(wy[0] * (-1.f - fy), sj[0], fmaf(wy[2] * (1.f - fy), sj[2], -fy * wy[1] * sj[1]));
}

// cold path of the backward pass: gathers the nine cotangent cells from the global image with the reference's index rule
template <bool WRAP>
__device__ __noinline__ float2 gather_fallback(const float* __restrict__ img, uint32_t xy, double2 th, double dt, int H, int W) {
    const Hit2 h = warp_hit2(xy, th, dt);
    if (!((fabs(h.xw) < 1.0e9) && (fabs(h.yw) < 1.0e9))) return make_float2(0.f, 0.f);
    if (h.rx + 1 < (WRAP ? -W : 0) || h.rx - 1 >= W || h.ry + 1 < (WRAP ? -H : 0) || h.ry - 1 >= H) return make_float2(0.f, 0.f);
    float d[9];
    for (int j = -1; j <= 1; ++j)
        for (int i = -1; i <= 1; ++i) {
            int rr = h.ry + j, cc = h.rx + i;
            d[(j + 1) * 3 + (i + 1)] = drop_index<WRAP>(rr, cc, H, W) ? __ldg(img + (int64_t)rr * W + cc) : 0.f;
        }
    float gx, gy;
    tap_gradient(d, h.fx, h.fy, gx, gy);
    return make_float2(gx, gy);
}

// ---- forward -----------------------------------------------------------------------------------------------------
// One CTA per chunk (grid-stride), RB reference times per pass with one window each.  Per pass: the windows follow from the
// theta range of the tile and the t range of the chunk (no pass over the events), are zeroed, voted into, and flushed.  The
// vote loop is branch-free: an event that misses its window (or a padding sentinel) votes zeros into a fixed cell of the
// window and is remembered in a bit mask; the rare misses are replayed through the out-of-line fallback afterwards.  Without
// branches the compiler interleaves the independent (event, reference time) pairs of a thread.
template <bool WRAP, int RB>
__device__ __forceinline__ void splat_tile_body(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, const Chunk* __restrict__ chunks,
             const float2* __restrict__ chunk_tr, const unsigned int* __restrict__ n_chunks_dev,
             const ThetaSrc& T, int H, int W, int R, const RefTimes& tref,
             const FixDst& dst /* [R][H*W] each */, int4* __restrict__ chunk_win /* [n_chunks][R] or null */,
             uint32_t* win /* [RB][kWinCap] dynamic shared memory */, const int* __restrict__ skip = nullptr) {
    __shared__ double2 th_s[kKeysPerTile];
    __shared__ int sbox[8][4];
    __shared__ Window swin[RB];
    __shared__ int4 srect[EINCM_MAX_REFS];
    __shared__ int s_off[EINCM_MAX_REFS + 1];
    __shared__ int s_sliced;
    const int64_t HW = (int64_t)H * W;
    const int tid = threadIdx.x;
    const int n_chunks = (int)__ldg(n_chunks_dev);
    const uint32_t th_base = smem_addr(th_s), win_base = smem_addr(win);
    const int Rpad = (R + RB - 1) / RB * RB;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const Chunk ch = chunks[c];
        const int ns = n_subs(ch);                   // sub-chunks of <= kSubChunk events: one in registers at a time
        EventGroup ev;
        int cur_sub = 0;
        const bool single = single_mode(ch);         // one event per thread (one sub-chunk)
        if (single) load_single_events<true>(ev_xy, ev_t, ch, ev); else load_sub_events<true>(ev_xy, ev_t, ch, 0, ev);
        // programmatic dependent launch: this kernel may have been scheduled while the previous kernel of the stream (the backward pass
        // of the previous evaluation: it clears the fixed-point images and reads chunk_win) was still running - only the staged
        // events were read so far
        asm volatile("griddepcontrol.wait;" ::: "memory");
        // unrolled solve graphs (eincm_minimize_bfgs_graph_host): the level ended in an earlier step of this launch - nothing to do
        if (skip != nullptr && *reinterpret_cast<const volatile int*>(skip) != 0) return;
        tile_theta_range(tile_theta(T, ch.origin, H, W, th_s), sbox);
        if (tid == 0) s_sliced = 0;
        int n_pairs = 0;                             // only used by the passes over further slices
        // misses are rare: the hot loop only counts hits; a thread whose count falls short repeats the window tests below
        int n_hit = 0, n_valid_ev = 0;               // over all sub-chunks of the thread
        bool counted = false;
        // One pass over the chunk for the windows of pairs p0 .. p0 + RB - 1.  PLAIN: pair j is reference time j (p0 < Rpad).
        auto pass = [&](auto plain_tag, const int p0) {
            constexpr bool PLAIN = decltype(plain_tag)::value;
            const int p_end = PLAIN ? R : n_pairs;
            // zero the cells in use
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int n4 = (swin[r].pw * swin[r].ph + 3) >> 2;
                for (int i = tid; i < n4; i += 256) reinterpret_cast<uint4*>(win + r * kWinCap)[i] = make_uint4(0u, 0u, 0u, 0u);
            }
            __syncthreads();
            // votes: every sub-chunk of the chunk into the same windows
            for (int sub = 0; sub < ns; ++sub) {
            if (sub != cur_sub) { load_sub_events<true>(ev_xy, ev_t, ch, sub, ev); cur_sub = sub; }
            if (!counted) {
#pragma unroll
                for (int k = 0; k < kEvK; ++k) n_valid_ev += ev.xy[k] != kNoEvent ? 1 : 0;
            }
            if (ev.xy[0] != kNoEvent) {              // groups are padded at their end only
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    if (p0 + r >= p_end) continue;
                    const Window wn = swin[r];
                    const uint32_t pitch4 = (uint32_t)wn.pw * 4u;
                    const uint32_t wb = win_base + (uint32_t)(r * kWinCap - (wn.oy * wn.pw + wn.ox)) * 4u;   // address of cell (0, 0) of the image
                    const double tr = PLAIN ? tref.t[p0 + r] : wn.tr;
                    // A thread's events are consecutive in the sorted stream (same source pixel, ascending time), so consecutive
                    // votes often share their centre cell: the nine tap sums stay pending in registers and go to shared memory
                    // only when the centre changes.  The shared-memory atomic pipe is the limiter of this kernel (ncu: l1tex ~73 %,
                    // issue ~45 %), so the extra selects are free and every merged vote saves nine atomic lanes.  A miss adds
                    // zeros to whatever is pending.
                    constexpr uint32_t kNone = 0xffffffffu;
                    int acc[9];
                    uint32_t acc_addr = kNone;
#pragma unroll
                    for (int q = 0; q < 9; ++q) acc[q] = 0;
#pragma unroll
                    for (int k = 0; k < kEvK; ++k) {
                        if (k == 1 && single) break;                             // CTA-uniform
                        const uint32_t xy = ev.xy[k];
                        const Hit2 h = warp_hit2(xy, lds_theta(th_base, xy), ev.t[k] - tr);
                        const bool valid = xy != kNoEvent;
                        const bool hit = valid & (fabs(h.xw - wn.cx) < wn.hx) & (fabs(h.yw - wn.cy) < wn.hy);
                        // a miss votes zeros: exp2(-(s * 1e4)^2) underflows to 0 for all three column factors
                        const TapsFix t = taps_fix(hit ? h.fx : 1.0e4f, hit ? h.fy : 0.f);
                        const uint32_t addr = hit ? wb + (uint32_t)(h.ry * wn.pw + h.rx) * 4u : acc_addr;
                        const bool same = addr == acc_addr;
                        if (!same && acc_addr != kNone) emit9(acc_addr, pitch4, acc);
#pragma unroll
                        for (int q = 0; q < 9; ++q) acc[q] = t.n[q] + (same ? acc[q] : 0);
                        acc_addr = addr;
                        n_hit += hit ? 1 : 0;
                    }
                    if (acc_addr != kNone) emit9(acc_addr, pitch4, acc);
                }
            }
            }
            counted = true;
            __syncthreads();
            // flush: non-zero window cells -> global fixed-point image (index rule applied here unless the window is interior).
            // Four cells per thread and round (one 128-bit load); most cells of a window are zero.
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                if (p0 + r >= p_end) continue;
                const Window wn = swin[r];
                const uint4* wr4 = reinterpret_cast<const uint4*>(win + r * kWinCap);
                const int64_t img_off = (int64_t)(PLAIN ? p0 + r : wn.img) * HW;
                const int cells = wn.pw * wn.ph;
                const bool interior = wn.interior != 0u;
                const uint32_t inv_pw = wn.inv_pw;
                unsigned long long* const img0 = dst.p[0] + img_off;
                for (int i4 = tid; 4 * i4 < cells; i4 += 256) {
                    const uint4 q = wr4[i4];
                    if ((q.x | q.y | q.z | q.w) == 0u) continue;
                    const uint32_t v4[4] = {q.x, q.y, q.z, q.w};
                    int row = (int)__umulhi((uint32_t)(4 * i4), inv_pw);
                    int col = 4 * i4 - row * wn.pw;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (v4[u] != 0u && 4 * i4 + u < cells) {
                            int rr = wn.oy + row, cc = wn.ox + col;
                            if (interior || drop_index<WRAP>(rr, cc, H, W)) {
                                atomicAdd(img0 + (rr * W + cc), (unsigned long long)v4[u]);
#pragma unroll 1
                                for (int p = 1; p < dst.n; ++p) atomicAdd(dst.p[p] + img_off + (rr * W + cc), (unsigned long long)v4[u]);
                            }
                        }
                        if (++col == wn.pw) { col = 0; ++row; }
                    }
                }
            }
        };
        auto count_valid = [&]() { return n_valid_ev * R; };      // event-reference pairs of the thread (after the first pass)
        // first slice of every reference time, RB reference times per pass (no rectangle is sliced in the common case)
        for (int r0 = 0; r0 < R; r0 += RB) {
            __syncthreads();                         // th_s / sbox ready (first pass); previous flush done
            if (tid < RB) {
                int4 rect = make_int4(0, 0, 0, 0);
                if (r0 + tid < R) {
                    int ax = sbox[0][0], bx = sbox[0][1], ay = sbox[0][2], by = sbox[0][3];
#pragma unroll
                    for (int w8 = 1; w8 < 8; ++w8) {
                        ax = min(ax, sbox[w8][0]); bx = max(bx, sbox[w8][1]); ay = min(ay, sbox[w8][2]); by = max(by, sbox[w8][3]);
                    }
                    const float2 tr = __ldg(chunk_tr + c);
                    const int x0 = (int)(ch.origin & 0xffffu), y0 = (int)(ch.origin >> 16);
                    rect = chunk_rect(x0, y0, min(x0 + kSortTile, W) - 1, min(y0 + kSortTile, H) - 1, ordered_float(ax), ordered_float(bx),
                                      ordered_float(ay), ordered_float(by), tr.x, tr.y, tref.t[r0 + tid]);
                    if (chunk_win != nullptr) chunk_win[(int64_t)c * R + r0 + tid] = rect;
                    srect[r0 + tid] = rect;
                    if (n_slices(rect) > 1) s_sliced = 1;
                }
                Window wn = slice_window(rect, 0);
                set_interior(wn, H, W);
                swin[tid] = wn;
            }
            __syncthreads();
            pass(std::true_type{}, r0);
        }
        // Further slices of sliced rectangles (large flows).  The rectangles are conservative bounds: when no event is left over
        // after the first slices, the further passes are skipped.  Any left-over event justifies them: the out-of-line fallback
        // of a thread is serial while the rest of the CTA waits (measured: tolerating 16 / 2 threads with left-over events per
        // pass made 50 .. 120 px flows 15 - 50 % / 10 - 20 % slower than always slicing, profiles/r1_large_flow.txt).
        int slices_done = kMaxSlices;
        if (s_sliced != 0) {
            const int n_thr = __syncthreads_count(n_hit != count_valid());      // also: flush done
            if (tid < 32) pair_offsets(tid < R ? n_slices(srect[tid]) : 0, R, Rpad, s_off);
            __syncthreads();
            n_pairs = s_off[R];
            if (n_thr > kMissThreadsPerPass * ((n_pairs - Rpad + RB - 1) / RB)) {
                for (int p0 = Rpad; p0 < n_pairs; p0 += RB) {
                    if (p0 > Rpad) __syncthreads();  // flush done
                    if (tid < RB) swin[tid] = pair_window(srect, s_off, R, Rpad, p0 + tid, tref, H, W);
                    __syncthreads();
                    pass(std::false_type{}, p0);
                }
            } else {
                slices_done = 1;
            }
        }
        if (n_hit != count_valid()) {                // events outside the (possibly cropped) rectangles, non-finite warps
#pragma unroll 1
            for (int sub = 0; sub < ns; ++sub) {
                if (sub != cur_sub) { load_sub_events<true>(ev_xy, ev_t, ch, sub, ev); cur_sub = sub; }
#pragma unroll
                for (int k = 0; k < kEvK; ++k) {
                    if (ev.xy[k] == kNoEvent) continue;
                    const double2 th = lds_theta(th_base, ev.xy[k]);
#pragma unroll 1
                    for (int r = 0; r < R; ++r) {
                        const double dt = ev.t[k] - tref.t[r];
                        const Hit2 h = warp_hit2(ev.xy[k], th, dt);
                        if (!hits_rect(srect[r], h.xw, h.yw, slices_done))
                            splat_fallback<WRAP>(dst, (int64_t)r * HW, ev.xy[k], th, dt, H, W);
                    }
                }
            }
        }
        __syncthreads();                             // th_s / sbox / swin / srect are rewritten for the next chunk
    }
}

template <bool WRAP, int RB>
__global__ void __launch_bounds__(256, 4)
k_splat_tile(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, const Chunk* __restrict__ chunks,
             const float2* __restrict__ chunk_tr, const unsigned int* __restrict__ n_chunks_dev,
             const ThetaSrc T, int H, int W, int R, const __grid_constant__ RefTimes tref,
             const __grid_constant__ FixDst dst /* [R][H*W] each */, int4* __restrict__ chunk_win /* [n_chunks][R] or null */,
             const int* __restrict__ skip /* non-zero: return at once (or null) */) {
    extern __shared__ __align__(16) uint32_t win[];          // [RB][kWinCap]
    splat_tile_body<WRAP, RB>(ev_xy, ev_t, chunks, chunk_tr, n_chunks_dev, T, H, W, R, tref, dst, chunk_win, win, skip);
}

// ---- batched form: one launch for B windows (blockIdx.y = window), one argument record per window in device memory ------------------
// BASELINE.json configs[2] / [3]: "batch of windows on 1 x B200".  The record of the CTA's window is copied into shared memory first.
template <typename Args>
__device__ __forceinline__ void load_args(Args& dst, const Args* __restrict__ src) {
    static_assert(sizeof(Args) % 4 == 0, "argument records are copied word by word");
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(&dst);
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < (int)(sizeof(Args) / 4); i += blockDim.x * blockDim.y) d[i] = __ldg(s + i);
    __syncthreads();
}

struct SplatArgs {
    const uint32_t* ev_xy; const double* ev_t; const Chunk* chunks; const float2* chunk_tr; const unsigned int* n_chunks_dev;
    ThetaSrc T; int H, W, R, pad; RefTimes tref; FixDst dst; int4* chunk_win;
    const int* skip;                // non-zero: the window's CTAs return at once (batched solve graphs), or null
};

template <bool WRAP, int RB>
__global__ void __launch_bounds__(256, 4)
k_splat_tile_b(const SplatArgs* __restrict__ args, const int* __restrict__ order) {
    extern __shared__ __align__(16) uint32_t win[];          // [RB][kWinCap]
    __shared__ SplatArgs sA;
    const int bw = batch_window(order);
    if (bw < 0) return;
    load_args(sA, args + bw);
    splat_tile_body<WRAP, RB>(sA.ev_xy, sA.ev_t, sA.chunks, sA.chunk_tr, sA.n_chunks_dev, sA.T, sA.H, sA.W, sA.R, sA.tref, sA.dst, sA.chunk_win, win, sA.skip);
}

// fixed-point image -> float64 image (paths that need the plain image: zero-warp image, event split, delta != 0)
__global__ void k_fix_to_f64(const unsigned long long* __restrict__ fix, int64_t n, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (double)(long long)fix[i] * kFixToIwe;
}

// ---- backward ----------------------------------------------------------------------------------------------------
// Same chunks and windows as the forward pass of the same theta (chunk_win).  dwin holds d loss / d IWE / (2 pi) (float32)
// of the window cells, zero where the index rule drops the cell.  Same branch-free structure as the forward pass: a miss
// reads a fixed cell and contributes zero, the rare misses are replayed through the out-of-line gather afterwards.
template <bool WRAP, int RB>
__device__ __forceinline__ void backward_tile_body(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, const Chunk* __restrict__ chunks, const unsigned int* __restrict__ n_chunks_dev,
                const ThetaSrc& T, int H, int W, int R, const RefTimes& tref,
                const float* __restrict__ dldi32 /* [R][H][W], scaled by 1/(2 pi) */, const int4* __restrict__ chunk_win,
                double* __restrict__ G /* [H][W][2] */, float* dwin /* [RB][kWinCap] dynamic shared memory */, const int* __restrict__ skip = nullptr) {
    __shared__ double2 th_s[kKeysPerTile];
    __shared__ Window swin[RB];
    __shared__ uint32_t c_xy[8], c_last[8];                  // single mode: first / last pixel of every warp's events ...
    __shared__ float c_sx[8], c_sy[8];                       // ... and the sums of the run that starts at lane 0
    const int64_t HW = (int64_t)H * W;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_chunks = (int)__ldg(n_chunks_dev);
    const uint32_t th_base = smem_addr(th_s), win_base = smem_addr(dwin);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const Chunk ch = chunks[c];
        const int ns = n_subs(ch);                   // sub-chunks of <= kSubChunk events: one in registers at a time
        EventGroup ev;
        const bool single = single_mode(ch);         // one event per thread (one sub-chunk)
        if (single) load_single_events<false>(ev_xy, ev_t, ch, ev); else load_sub_events<false>(ev_xy, ev_t, ch, 0, ev);
        tile_theta(T, ch.origin, H, W, th_s);
        // programmatic dependent launch: everything above read the staged events and the flow operand only
        if (c == (int)blockIdx.x) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            if (skip != nullptr && *reinterpret_cast<const volatile int*>(skip) != 0) return;
        }
        // With R <= RB (one group of reference times: the shipped recipes up to R = 4) the windows are filled once and every sub-chunk
        // gathers from them; otherwise they are refilled per (sub-chunk, group).
        const bool one_group = R <= RB;
        bool filled = false;
        for (int sub = 0; sub < ns; ++sub) {
        if (sub > 0) load_sub_events<false>(ev_xy, ev_t, ch, sub, ev);
        const bool active = ev.xy[0] != kNoEvent;    // groups are padded at their end only
        float ax[kEvK], ay[kEvK];
#pragma unroll
        for (int k = 0; k < kEvK; ++k) { ax[k] = 0.f; ay[k] = 0.f; }
        for (int r0 = 0; r0 < R; r0 += RB) {
            if (!(one_group && filled)) {
            if (filled) __syncthreads();             // previous readers of swin / dwin are done
            filled = true;
            if (tid < RB) {
                int4 q = make_int4(0, 0, 0, 0);
                if (r0 + tid < R) q = chunk_win[(int64_t)c * R + r0 + tid];
                Window wn = slice_window(q, 0);      // a sliced rectangle (large flow): first slice here, the rest gathers from the global image
                set_interior(wn, H, W);
                swin[tid] = wn;
            }
            __syncthreads();
            // window cells <- d loss / d IWE: all global loads of a round (four cells per thread) are issued before the stores
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                if (r0 + r >= R) continue;
                const Window wn = swin[r];
                const float* img = dldi32 + (int64_t)(r0 + r) * HW;
                float* wr = dwin + r * kWinCap;
                const int cells = wn.pw * wn.ph;
                const bool interior = wn.interior != 0u;
                const uint32_t inv_pw = wn.inv_pw;
                for (int i0 = tid; i0 < cells; i0 += 4 * 256) {
                    float v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * 256;
                        const int row = (int)__umulhi((uint32_t)min(i, cells - 1), inv_pw), col = min(i, cells - 1) - row * wn.pw;
                        int rr = wn.oy + row, cc = wn.ox + col;
                        v[u] = (i < cells && (interior || drop_index<WRAP>(rr, cc, H, W))) ? __ldg(img + (rr * W + cc)) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * 256;
                        if (i < cells) wr[i] = v[u];
                    }
                }
            }
            __syncthreads();
            }
            int n_hit = 0, n_valid = 0;
            if (active) {
#pragma unroll
                for (int k = 0; k < kEvK; ++k) n_valid += ev.xy[k] != kNoEvent ? 1 : 0;
                n_valid *= min(RB, R - r0);
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    if (r0 + r >= R) continue;
                    const Window wn = swin[r];
                    const uint32_t pitch4 = (uint32_t)wn.pw * 4u;
                    const uint32_t wb = win_base + (uint32_t)(r * kWinCap - (wn.oy * wn.pw + wn.ox)) * 4u;
                    const uint32_t safe = win_base + (uint32_t)(r * kWinCap) * 4u + pitch4 + 4u;
                    const double tr = tref.t[r0 + r];
#pragma unroll
                    for (int k = 0; k < kEvK; ++k) {
                        if (k == 1 && single) break;                             // CTA-uniform
                        const uint32_t xy = ev.xy[k];
                        const double dt = ev.t[k] - tr;
                        const Hit2 h = warp_hit2(xy, lds_theta(th_base, xy), dt);
                        const bool valid = xy != kNoEvent;
                        const bool hit = valid & (fabs(h.xw - wn.cx) < wn.hx) & (fabs(h.yw - wn.cy) < wn.hy);
                        const uint32_t mid = hit ? wb + (uint32_t)(h.ry * wn.pw + h.rx) * 4u : safe;
                        const uint32_t up = mid - pitch4, dn = mid + pitch4;
                        float d[9];
                        asm volatile("ld.shared.f32 %0, [%9 + -4];\n\tld.shared.f32 %1, [%9];\n\tld.shared.f32 %2, [%9 + 4];\n\t"
                                     "ld.shared.f32 %3, [%10 + -4];\n\tld.shared.f32 %4, [%10];\n\tld.shared.f32 %5, [%10 + 4];\n\t"
                                     "ld.shared.f32 %6, [%11 + -4];\n\tld.shared.f32 %7, [%11];\n\tld.shared.f32 %8, [%11 + 4];"
                                     : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]), "=f"(d[4]), "=f"(d[5]), "=f"(d[6]), "=f"(d[7]), "=f"(d[8])
                                     : "r"(up), "r"(mid), "r"(dn));
                        float gx, gy;
                        tap_gradient(d, hit ? h.fx : 0.f, hit ? h.fy : 0.f, gx, gy);
                        const float ndt = -(float)dt;
                        // a miss may have read anything (and a padding sentinel has no timestamp): select, do not multiply by zero
                        ax[k] = hit ? fmaf(ndt, gx, ax[k]) : ax[k];
                        ay[k] = hit ? fmaf(ndt, gy, ay[k]) : ay[k];
                        n_hit += hit ? 1 : 0;
                    }
                }
            }
            if (n_hit != n_valid) {
#pragma unroll
                for (int k = 0; k < kEvK; ++k) {
                    if (ev.xy[k] == kNoEvent) continue;
                    const double2 th = lds_theta(th_base, ev.xy[k]);
#pragma unroll 1
                    for (int r = 0; r < RB && r0 + r < R; ++r) {
                        const double dt = ev.t[k] - tref.t[r0 + r];
                        const Hit2 h = warp_hit2(ev.xy[k], th, dt);
                        const Window& wn = swin[r];
                        if (!((fabs(h.xw - wn.cx) < wn.hx) & (fabs(h.yw - wn.cy) < wn.hy))) {
                            const float2 g = gather_fallback<WRAP>(dldi32 + (int64_t)(r0 + r) * HW, ev.xy[k], th, dt, H, W);
                            ax[k] = fmaf(-(float)dt, g.x, ax[k]);
                            ay[k] = fmaf(-(float)dt, g.y, ay[k]);
                        }
                    }
                }
            }
        }
        // per-thread runs of equal source pixel: all but the last run go straight to G
        uint32_t run_xy = ev.xy[0];
        float sx = ax[0], sy = ay[0];
        if (!single) {
#pragma unroll
            for (int k = 1; k < kEvK; ++k) {
                if (ev.xy[k] == run_xy) { sx += ax[k]; sy += ay[k]; }
                else {
                    if (run_xy != kNoEvent) red_G(G, W, run_xy, sx, sy);
                    run_xy = ev.xy[k]; sx = ax[k]; sy = ay[k];
                }
            }
        }
        // last runs of the warp's threads: segmented (by source pixel) suffix sum, one reduction pair per run
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float ox = __shfl_down_sync(0xffffffffu, sx, o);
            const float oy = __shfl_down_sync(0xffffffffu, sy, o);
            const uint32_t oxy = __shfl_down_sync(0xffffffffu, run_xy, o);
            if (lane + o < 32 && oxy == run_xy) { sx += ox; sy += oy; }
        }
        const uint32_t prev = __shfl_up_sync(0xffffffffu, run_xy, 1);
        if (single) {
            // one event per thread, sorted by pixel: a pixel's run may cross warps.  The run's first warp collects the sums of the warps
            // it continues into (fixed order) and issues the ONE reduction pair of the pixel: the gradient of a window whose tiles hold a
            // single chunk each does not depend on the order of any atomics.
            const int wid = tid >> 5;
            if (lane == 0) { c_xy[wid] = run_xy; c_sx[wid] = sx; c_sy[wid] = sy; }
            if (lane == 31) c_last[wid] = run_xy;
            __syncthreads();
            const bool head = (lane == 0) ? (wid == 0 || c_last[wid - 1] != run_xy) : (prev != run_xy);
            if (head && run_xy != kNoEvent) {
                if (c_last[wid] == run_xy) {                        // the run reaches the end of this warp
                    for (int q = wid + 1; q < 8 && c_xy[q] == run_xy; ++q) {
                        sx += c_sx[q]; sy += c_sy[q];
                        if (c_last[q] != run_xy) break;
                    }
                }
                red_G(G, W, run_xy, sx, sy);
            }
        } else {
            const bool head = (lane == 0) || (prev != run_xy);
            if (head && run_xy != kNoEvent) red_G(G, W, run_xy, sx, sy);
        }
        }
        __syncthreads();                             // th_s / swin / dwin are rewritten for the next chunk
    }
}

template <bool WRAP, int RB>
__global__ void __launch_bounds__(256, 4)
k_backward_tile(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, const Chunk* __restrict__ chunks, const unsigned int* __restrict__ n_chunks_dev,
                const ThetaSrc T, int H, int W, int R, const __grid_constant__ RefTimes tref,
                const float* __restrict__ dldi32 /* [R][H][W], scaled by 1/(2 pi) */, const int4* __restrict__ chunk_win,
                double* __restrict__ G /* [H][W][2] */, const int* __restrict__ skip /* non-zero: return at once (or null) */) {
    extern __shared__ __align__(16) float dwin[];            // [RB][kWinCap]
    backward_tile_body<WRAP, RB>(ev_xy, ev_t, chunks, n_chunks_dev, T, H, W, R, tref, dldi32, chunk_win, G, dwin, skip);
}

struct BackwardTileArgs {
    const uint32_t* ev_xy; const double* ev_t; const Chunk* chunks; const unsigned int* n_chunks_dev;
    ThetaSrc T; int H, W, R, pad; RefTimes tref; const float* dldi32; const int4* chunk_win; double* G;
    const int* skip;                // as in SplatArgs
};

template <bool WRAP, int RB>
__global__ void __launch_bounds__(256, 4)
k_backward_tile_b(const BackwardTileArgs* __restrict__ args, const int* __restrict__ order) {
    extern __shared__ __align__(16) float dwin[];            // [RB][kWinCap]
    __shared__ BackwardTileArgs sA;
    const int bw = batch_window(order);
    if (bw < 0) return;
    load_args(sA, args + bw);
    backward_tile_body<WRAP, RB>(sA.ev_xy, sA.ev_t, sA.chunks, sA.n_chunks_dev, sA.T, sA.H, sA.W, sA.R, sA.tref, sA.dldi32, sA.chunk_win, sA.G, dwin, sA.skip);
}

// ---- debug tap: Xs_rounded (event_utils.py:33) of reference r, written back in the original event order ----
// EXACT selects the conversion used by the float64 nine-tap kernels (cvt.rni), otherwise the magic-constant rint of the tile
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
            const Hit h = warp_hit((double)x, (double)y, th.x, th.y, ev_t[e] - t_ref);
            cols_out[o] = h.ok ? h.rx : INT32_MAX;
            rows_out[o] = h.ok ? h.ry : INT32_MAX;
        }
    }
}

}  // namespace eincm
