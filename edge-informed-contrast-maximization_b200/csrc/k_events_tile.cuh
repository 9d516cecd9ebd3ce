// Tile-privatised event-space kernels (default path).
//
// The staged event stream is sorted by source pixel, tile-major over 16x16 source tiles, and by time inside a pixel
// (k_prep.cuh); every tile's segment is padded to a multiple of kEvK events and cut into chunks of <= kChunk events.
// One CTA owns one chunk.  Because all events of a chunk start inside one 16x16 tile, their warped 3x3 patches
// (reference src/utils/event_utils.py:41-59) fall into a small rectangle of the destination image per reference time:
// the CTA measures that rectangle (pre-pass: warp + rint only), keeps it as a window in shared memory, votes into it
// with native 32-bit shared-memory integer atomics (ATOMS.ADD), and finally adds the non-zero window cells to the
// global image with one 64-bit integer reduction each.  Votes are fixed point, 2^kFixShift * (2 pi v):
//   * shared-memory float atomics are CAS loops on sm_100a, integer adds are native (see profiles/microbench);
//   * integer sums do not depend on the order of the votes, so the image of warped events - and therefore the
//     objective - is bit-reproducible from run to run, which matters to BFGS with gtol = 1e-7 (main.yaml:34).
// A vote is quantised to 2^-21 of the centre-tap value (4.8e-7, about 4 float32 ulps of the largest tap); coordinates,
// the warp and rint() stay float64, so pixel indices are bit-exact.  Events whose patch leaves the window (bounding
// rectangle larger than kWinCap cells: very large flow) fall back to per-tap global reductions.  The reference's index
// rule (negative indices wrap, out of range drops: SURVEY.md A.4) is applied once per window cell at flush time.
//
// Backward: the same windows (recorded by the forward pass) are filled with d loss / d IWE and the nine taps of every
// event are read from shared memory instead of global memory.
#pragma once
#include <climits>

#include "common.cuh"
#include "k_events.cuh"
#include "k_prep.cuh"
#include "k_theta.cuh"

namespace eincm {

constexpr int kEvK = 4;                    // events per thread (one 128-bit load of packed coordinates, two of timestamps)
constexpr uint32_t kNoEvent = 0xffffffffu; // padding sentinel of the sorted stream
constexpr int kMaxRB = 4;                  // reference times processed per pass over a chunk

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


constexpr int kChunk = (int)kChunkEvents;   // events per chunk = kEvK events per thread x 256 threads
static_assert(kChunk == kEvK * 256 && kStreamAlign == kEvK, "one chunk = one CTA pass of kEvK events per thread");
constexpr int kWinCap = 4096;         // window cells per reference time (16 KB of uint32 / float)
constexpr int kWinMaxH = 64;          // rows kept when the bounding rectangle exceeds kWinCap
constexpr int kFixShift = 21;
constexpr float kFixScale = 2097152.0f;                  // 2^21
constexpr double kFixToIwe = kInv2Pi / 2097152.0;        // fixed-point sum -> image value
constexpr float kRoundMagic = 12582912.0f;               // 1.5 * 2^23: float -> int by mantissa alignment
constexpr int kRoundMagicBits = 0x4B400000;

struct TapsFix { int n[9]; };                            // index (j+1)*3 + (i+1): column offset i, row offset j

// exp(-0.5 (d - f)^2) for d = -1, 0, 1:  u = s f, q_d = s d - u, value = 2^(-q_d^2) with s = sqrt(0.5 log2 e)
__device__ __forceinline__ void axis_exp3(float f, float e[3]) {
    constexpr float s = 0.84932180028801907f;
    const float u = s * f;
    const float q0 = -s - u, q2 = s - u;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[0]) : "f"(-q0 * q0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[1]) : "f"(-u * u));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[2]) : "f"(-q2 * q2));
}

// nine fixed-point tap values of one warped event: round(2^21 * exp(-0.5 ((i - fx)^2 + (j - fy)^2)))
__device__ __forceinline__ TapsFix taps_fix(float fx, float fy) {
    float ex[3], ey[3];
    axis_exp3(fx, ex);
    axis_exp3(fy, ey);
#pragma unroll
    for (int d = 0; d < 3; ++d) ey[d] *= kFixScale;
    TapsFix t;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) t.n[j * 3 + i] = __float_as_int(fmaf(ex[i], ey[j], kRoundMagic)) - kRoundMagicBits;
    return t;
}

__device__ __forceinline__ void load_chunk_events(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, const Chunk ch,
                                                  EventGroup& ev) {
    if (4u * threadIdx.x < ch.count) {
        load_group(ev_xy, ev_t, (int64_t)(ch.start >> 2) + threadIdx.x, ev);
    } else {
#pragma unroll
        for (int k = 0; k < kEvK; ++k) { ev.xy[k] = kNoEvent; ev.t[k] = 0.0; }
    }
}

// Window of one reference time inside the destination image: origin (ox, oy), pw x ph cells, row pitch pw, and the
// multiplier of the exact division i / pw for i < 2^16 (cell index -> row, column without an integer division).
struct Window { int ox, oy, pw, ph; uint32_t inv_pw; };

__device__ __forceinline__ Window make_window(int mnx, int mny, int mxx, int mxy) {
    Window w{0, 0, 0, 0, 0u};
    if (mnx > mxx || mny > mxy) return w;                    // no valid event
    long long bw = (long long)mxx - mnx + 3, bh = (long long)mxy - mny + 3;
    if (bw * bh > kWinCap) {
        if (bh > kWinMaxH) bh = kWinMaxH;
        if (bw > kWinCap / bh) bw = kWinCap / bh;
    }
    w.ox = mnx - 1; w.oy = mny - 1; w.pw = (int)bw; w.ph = (int)bh;
    w.inv_pw = 0xffffffffu / (uint32_t)bw + 1u;              // ceil(2^32 / pw): exact cell -> row for cells < 2^16
    return w;
}

__device__ __forceinline__ void cell_to_rc(const Window& w, int i, int& row, int& col) {
    row = (int)__umulhi((uint32_t)i, w.inv_pw);
    col = i - row * w.pw;
}

// the patch of a warped event with rounded centre (rx, ry) lies inside the window
__device__ __forceinline__ bool in_window(const Window& w, int rx, int ry) {
    return ((unsigned)(rx - w.ox - 1) < (unsigned)(w.pw - 2)) & ((unsigned)(ry - w.oy - 1) < (unsigned)(w.ph - 2));
}

// Adds the nine pending tap sums to the 3x3 cells around the shared-memory cell `mid` (rows `pitch4` bytes apart): nine
// `red.shared.add.u32` (ATOMS.ADD without return value) with immediate column offsets.  Callers branch around the whole
// block: ptxas turns a predicated shared-memory atomic into its own branch, which costs four instructions per vote.
__device__ __forceinline__ void emit9(uint32_t mid, uint32_t pitch4, const int (&a)[9]) {
    const uint32_t up = mid - pitch4, dn = mid + pitch4;
    asm volatile(
        "red.shared.add.u32 [%0 + -4], %3;\n\tred.shared.add.u32 [%0], %4;\n\tred.shared.add.u32 [%0 + 4], %5;\n\t"
        "red.shared.add.u32 [%1 + -4], %6;\n\tred.shared.add.u32 [%1], %7;\n\tred.shared.add.u32 [%1 + 4], %8;\n\t"
        "red.shared.add.u32 [%2 + -4], %9;\n\tred.shared.add.u32 [%2], %10;\n\tred.shared.add.u32 [%2 + 4], %11;"
        ::"r"(up), "r"(mid), "r"(dn), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8])
        : "memory");
}

// theta of the 16x16 source tile at `origin` -> shared memory (zero flow when T.theta is null)
__device__ __forceinline__ void tile_theta(const ThetaSrc& T, uint32_t origin, int H, int W, double2* __restrict__ th_s) {
    const int p = threadIdx.x;                       // 256 threads <-> 256 pixels
    const int x = (int)(origin & 0xffffu) + (p & 15), y = (int)(origin >> 16) + (p >> 4);
    double2 v = make_double2(0.0, 0.0);
    if (T.theta != nullptr && x < W && y < H) v = theta_at(T, x, y);
    th_s[p] = v;
}

__device__ __forceinline__ double2 event_theta(const double2* __restrict__ th_s, uint32_t xy) {
    return th_s[((xy >> 12) & 0xf0u) | (xy & 0xfu)];         // (y & 15) * 16 + (x & 15)
}

// Destination images of the flush.  Single GPU: the plan's own fixed-point image.  Event split with peer access: the images
// of ALL ranks of the split (own + peers' device pointers opened through CUDA IPC) - every rank adds its non-zero window
// cells to every rank's image with 64-bit integer reductions over NVLink, so the "all-reduce of the partial images" is fused
// into the splat: integer sums are order-independent, every rank ends up with the identical complete image.
constexpr int kMaxPeers = 8;
struct FixDst { unsigned long long* p[kMaxPeers]; int n; };

// cold path of the splat: an event whose patch leaves its window adds its taps to the global images one by one, with the
// reference's index rule (out of line: keeps the hot loop small - the kernel was instruction-cache bound with it inlined)
template <bool WRAP>
__device__ __noinline__ void splat_fallback(const FixDst& dst, int64_t img_off, int rx, int ry, const TapsFix& t, int H, int W) {
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) {
            int rr = ry + j - 1, cc = rx + i - 1;
            if (drop_index<WRAP>(rr, cc, H, W)) {
                const int64_t off = img_off + (int64_t)rr * W + cc;
#pragma unroll 1
                for (int q = 0; q < dst.n; ++q) atomicAdd(dst.p[q] + off, (unsigned long long)t.n[j * 3 + i]);
            }
        }
}

template <bool WRAP>
__device__ __noinline__ void gather_fallback(const float* __restrict__ img, int rx, int ry, int H, int W, float (&d)[9]) {
    for (int j = -1; j <= 1; ++j)
        for (int i = -1; i <= 1; ++i) {
            int rr = ry + j, cc = rx + i;
            d[(j + 1) * 3 + (i + 1)] = drop_index<WRAP>(rr, cc, H, W) ? __ldg(img + (int64_t)rr * W + cc) : 0.f;
        }
}

// ---- forward -----------------------------------------------------------------------------------------------------
// One CTA per chunk (grid-stride), RB reference times per pass with one window each.  Per pass: zero the windows, measure
// the bounding rectangles (float32 pre-pass), vote, flush.  A thread walks its kEvK consecutive events per reference time and
// merges a vote into the next one when both have the same centre cell (events of one source pixel are time-sorted, so this is
// common): the nine pending tap sums stay in registers and are only sent to shared memory when the centre changes.  The merge
// is branch-free (predicated adds / predicated reductions), so diverging lanes cost nothing extra.
template <bool WRAP, int RB>
__global__ void __launch_bounds__(256, 4)
k_splat_tile(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, const Chunk* __restrict__ chunks, const unsigned int* __restrict__ n_chunks_dev,
             const ThetaSrc T, int H, int W, int R, const __grid_constant__ RefTimes tref,
             const __grid_constant__ FixDst dst /* [R][H*W] each */, int4* __restrict__ chunk_win /* [n_chunks][R] or null */) {
    extern __shared__ __align__(16) uint32_t win[];          // [RB][kWinCap]
    __shared__ double2 th_s[kKeysPerTile];
    __shared__ int sbox[8][RB][4];
    __shared__ Window swin[RB];
    const int64_t HW = (int64_t)H * W;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n_chunks = (int)__ldg(n_chunks_dev);
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const Chunk ch = chunks[c];
        EventGroup ev;
        load_chunk_events(ev_xy, ev_t, ch, ev);
        tile_theta(T, ch.origin, H, W, th_s);
        for (int r0 = 0; r0 < R; r0 += RB) {
            // zero the windows (whole capacity: a handful of 128-bit stores per thread)
            for (int i = tid; i < RB * kWinCap / 4; i += 256) reinterpret_cast<uint4*>(win)[i] = make_uint4(0u, 0u, 0u, 0u);
            __syncthreads();                         // th_s ready (first pass); previous flush done
            // pre-pass: bounding rectangle of the rounded warped pixels per reference time.  float32 arithmetic (error far
            // below the 0.01 px margin for any flow a window can hold); an event the rectangle misses takes the fallback.
            {
                float lox[RB], loy[RB], hix[RB], hiy[RB], trf[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    lox[r] = 3.0e9f; loy[r] = 3.0e9f; hix[r] = -3.0e9f; hiy[r] = -3.0e9f;
                    trf[r] = (float)tref.t[min(r0 + r, EINCM_MAX_REFS - 1)];
                }
#pragma unroll
                for (int k = 0; k < kEvK; ++k) {
                    if (ev.xy[k] == kNoEvent) continue;
                    const double2 th = event_theta(th_s, ev.xy[k]);
                    const float xf = (float)(ev.xy[k] & 0xffffu), yf = (float)(ev.xy[k] >> 16), tf = (float)ev.t[k];
                    const float thx = (float)th.x, thy = (float)th.y;
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        const float dt = tf - trf[r];
                        const float xw = fmaf(-thx, dt, xf), yw = fmaf(-thy, dt, yf);
                        lox[r] = fminf(lox[r], xw); hix[r] = fmaxf(hix[r], xw);
                        loy[r] = fminf(loy[r], yw); hiy[r] = fmaxf(hiy[r], yw);
                    }
                }
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    // clamp (huge / non-finite flows end up in the fallback path anyway), widen by the rounding margin
                    const int a = __float2int_rd(fmaxf(lox[r], -1.0e9f) - 0.51f), b = __float2int_rd(fmaxf(loy[r], -1.0e9f) - 0.51f);
                    const int cx = __float2int_ru(fminf(hix[r], 1.0e9f) + 0.51f), d = __float2int_ru(fminf(hiy[r], 1.0e9f) + 0.51f);
                    const int ra = __reduce_min_sync(0xffffffffu, a), rb = __reduce_min_sync(0xffffffffu, b);
                    const int rc = __reduce_max_sync(0xffffffffu, cx), rd = __reduce_max_sync(0xffffffffu, d);
                    if (lane == 0) { sbox[wid][r][0] = ra; sbox[wid][r][1] = rb; sbox[wid][r][2] = rc; sbox[wid][r][3] = rd; }
                }
            }
            __syncthreads();
            if (tid < RB) {
                int a = INT_MAX, b = INT_MAX, cmx = INT_MIN, d = INT_MIN;
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) {
                    a = min(a, sbox[w8][tid][0]); b = min(b, sbox[w8][tid][1]);
                    cmx = max(cmx, sbox[w8][tid][2]); d = max(d, sbox[w8][tid][3]);
                }
                const Window wn = make_window(a, b, cmx, d);
                swin[tid] = wn;
                if (chunk_win != nullptr && r0 + tid < R) chunk_win[(int64_t)c * R + r0 + tid] = make_int4(wn.ox, wn.oy, wn.pw, wn.ph);
            }
            __syncthreads();
            // main pass
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                if (r0 + r >= R) continue;
                const Window wn = swin[r];
                const uint32_t wbase = (uint32_t)__cvta_generic_to_shared(win + r * kWinCap);
                const uint32_t pitch4 = (uint32_t)wn.pw * 4u;
                const double tr = tref.t[r0 + r];
                int acc[9];
                uint32_t acc_addr = 0u;               // shared address of the pending centre cell; 0 = nothing pending
#pragma unroll
                for (int q = 0; q < 9; ++q) acc[q] = 0;
#pragma unroll
                for (int k = 0; k < kEvK; ++k) {
                    if (ev.xy[k] == kNoEvent) continue;
                    const double2 th = event_theta(th_s, ev.xy[k]);
                    const double xd = (double)(ev.xy[k] & 0xffffu), yd = (double)(ev.xy[k] >> 16);
                    const Hit h = warp_hit(xd, yd, th.x, th.y, ev.t[k] - tr);
                    if (!h.ok) continue;
                    const TapsFix t = taps_fix(h.fx, h.fy);
                    if (in_window(wn, h.rx, h.ry)) {
                        const uint32_t addr = wbase + (uint32_t)((h.ry - wn.oy) * wn.pw + (h.rx - wn.ox)) * 4u;
                        const bool same = addr == acc_addr;
                        const bool emit = !same && acc_addr != 0u;
                        // send the pending sums (other centre), then fold them into the new taps when the centre is the same
                        if (emit) emit9(acc_addr, pitch4, acc);
#pragma unroll
                        for (int q = 0; q < 9; ++q) acc[q] = t.n[q] + (same ? acc[q] : 0);
                        acc_addr = addr;
                    } else {
                        splat_fallback<WRAP>(dst, (int64_t)(r0 + r) * HW, h.rx, h.ry, t, H, W);
                    }
                }
                if (acc_addr != 0u) emit9(acc_addr, pitch4, acc);
            }
            __syncthreads();
            // flush: non-zero window cells -> global fixed-point image (index rule applied here unless the window is interior)
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                if (r0 + r >= R) continue;
                const Window wn = swin[r];
                const uint32_t* wr = win + r * kWinCap;
                const int64_t img_off = (int64_t)(r0 + r) * HW;
                const int cells = wn.pw * wn.ph;
                const bool interior = wn.ox >= 0 && wn.oy >= 0 && wn.ox + wn.pw <= W && wn.oy + wn.ph <= H;
                for (int i = tid; i < cells; i += 256) {
                    const uint32_t v = wr[i];
                    if (v != 0u) {
                        int row, col;
                        cell_to_rc(wn, i, row, col);
                        int rr = wn.oy + row, cc = wn.ox + col;
                        if (interior || drop_index<WRAP>(rr, cc, H, W)) {
                            const int64_t off = img_off + (rr * W + cc);
#pragma unroll 1
                            for (int q = 0; q < dst.n; ++q) atomicAdd(dst.p[q] + off, (unsigned long long)v);
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
}

// fixed-point image -> float64 image (paths that need the plain image: zero-warp image, event split, delta != 0)
__global__ void k_fix_to_f64(const unsigned long long* __restrict__ fix, int64_t n, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (double)(long long)fix[i] * kFixToIwe;
}

// ---- backward ----------------------------------------------------------------------------------------------------
// Same chunks and windows as the forward pass of the same theta (chunk_win).  dwin holds d loss / d IWE / (2 pi) (float32)
// of the window cells, zero where the index rule drops the cell.
template <bool WRAP, int RB>
__global__ void __launch_bounds__(256, 4)
k_backward_tile(const uint32_t* __restrict__ ev_xy, const double* __restrict__ ev_t, const Chunk* __restrict__ chunks, const unsigned int* __restrict__ n_chunks_dev,
                const ThetaSrc T, int H, int W, int R, const __grid_constant__ RefTimes tref,
                const float* __restrict__ dldi32 /* [R][H][W], scaled by 1/(2 pi) */, const int4* __restrict__ chunk_win,
                double* __restrict__ G /* [H][W][2] */) {
    extern __shared__ __align__(16) float dwin[];            // [RB][kWinCap]
    __shared__ double2 th_s[kKeysPerTile];
    __shared__ Window swin[RB];
    const int64_t HW = (int64_t)H * W;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_chunks = (int)__ldg(n_chunks_dev);
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const Chunk ch = chunks[c];
        EventGroup ev;
        load_chunk_events(ev_xy, ev_t, ch, ev);
        tile_theta(T, ch.origin, H, W, th_s);
        float ax[kEvK], ay[kEvK];
#pragma unroll
        for (int k = 0; k < kEvK; ++k) { ax[k] = 0.f; ay[k] = 0.f; }
        for (int r0 = 0; r0 < R; r0 += RB) {
            if (tid < RB && r0 + tid < R) {
                const int4 q = chunk_win[(int64_t)c * R + r0 + tid];
                Window wn{q.x, q.y, q.z, q.w, 0u};
                if (q.z > 0) wn.inv_pw = 0xffffffffu / (uint32_t)q.z + 1u;
                swin[tid] = wn;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                if (r0 + r >= R) continue;
                const Window wn = swin[r];
                const float* img = dldi32 + (int64_t)(r0 + r) * HW;
                float* wr = dwin + r * kWinCap;
                const int cells = wn.pw * wn.ph;
                const bool interior = wn.ox >= 0 && wn.oy >= 0 && wn.ox + wn.pw <= W && wn.oy + wn.ph <= H;
                for (int i0 = tid; i0 < cells; i0 += 4 * 256) {       // all global loads of a round first, then the stores
                    float v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * 256;
                        int row, col;
                        cell_to_rc(wn, min(i, cells - 1), row, col);
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
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                if (r0 + r >= R) continue;
                const Window wn = swin[r];
                const float* wr = dwin + r * kWinCap;
                const double tr = tref.t[r0 + r];
                const float trf = (float)tr;
#pragma unroll
                for (int k = 0; k < kEvK; ++k) {
                    if (ev.xy[k] == kNoEvent) continue;
                    const double2 th = event_theta(th_s, ev.xy[k]);
                    const double xd = (double)(ev.xy[k] & 0xffffu), yd = (double)(ev.xy[k] >> 16);
                    const Hit h = warp_hit(xd, yd, th.x, th.y, ev.t[k] - tr);
                    if (!h.ok) continue;
                    float d[9];
                    if (in_window(wn, h.rx, h.ry)) {
                        const float* p = wr + (h.ry - wn.oy) * wn.pw + (h.rx - wn.ox);
#pragma unroll
                        for (int j = -1; j <= 1; ++j)
#pragma unroll
                            for (int i = -1; i <= 1; ++i) d[(j + 1) * 3 + (i + 1)] = p[j * wn.pw + i];
                    } else {
                        gather_fallback<WRAP>(dldi32 + (int64_t)(r0 + r) * HW, h.rx, h.ry, H, W, d);
                    }
                    // separable evaluation: wx_i = exp(-0.5 (i - fx)^2), s_j = sum_i D_ij wx_i, sx_j = sum_i D_ij wx_i (i - fx)
                    //   dL/dx' = sum_j wy_j sx_j,   dL/dy' = sum_j wy_j (j - fy) s_j        (D already carries 1/(2 pi))
                    const float fx = h.fx, fy = h.fy;
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
                    const float gx = fmaf(wy[0], sxj[0], fmaf(wy[2], sxj[2], wy[1] * sxj[1]));
                    const float gy = fmaf(wy[0] * (-1.f - fy), sj[0], fmaf(wy[2] * (1.f - fy), sj[2], -fy * wy[1] * sj[1]));
                    const float dtf = (float)ev.t[k] - trf;
                    ax[k] = fmaf(-dtf, gx, ax[k]);
                    ay[k] = fmaf(-dtf, gy, ay[k]);
                }
            }
            __syncthreads();
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
        // last runs of the warp's threads: segmented (by source pixel) suffix sum, one reduction pair per run
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
