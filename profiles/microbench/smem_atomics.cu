// Microbenchmark: shared-memory accumulation throughput on B200 for the 3x3 vote pattern of the EINCM splat.
// Each thread handles `iters` pseudo-events; an event adds 9 values to the 3x3 neighbourhood of a pixel of a PWxPW window
// held in shared memory.  Variants:
//   u32   : 9 x ATOMS.ADD (native 32-bit integer, what a fixed-point accumulator would use)
//   f32   : 9 x atomicAdd(float) on shared memory (compiles to an ATOMS.CAST.SPIN loop)
//   lds   : 9 x LDS gather (the backward pass pattern), summed in registers
//   sts   : 9 x plain STS (not a valid accumulation; the LSU floor for reference)
// Address patterns: `spread` = every lane an independent pseudo-random pixel; `pairs` = lanes 2k, 2k+1 hit the same pixel
// (events sorted by pixel and time often share their destination); `line` = a warp's 32 events walk along a short line.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int PW = 48;
constexpr int NPLANE = 3;

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int PATTERN>
__device__ __forceinline__ int pixel_of(uint32_t i, int lane) {
    uint32_t key = i;
    if (PATTERN == 1) key = i >> 1;
    if (PATTERN == 2) {
        // the warp's events walk along a line of ~16 px through a warp-specific anchor
        const uint32_t h = hash32(i >> 5);
        const int ax = 4 + (int)(h % 24), ay = 4 + (int)((h >> 8) % 36);
        const int step = lane >> 1;
        return (ay + (step >> 2)) * PW + ax + step;
    }
    const uint32_t h = hash32(key);
    const int x = 1 + (int)(h % (PW - 2)), y = 1 + (int)((h >> 12) % (PW - 2));
    return y * PW + x;
}

template <int MODE, int PATTERN>
__global__ void __launch_bounds__(256) k(uint32_t* out, int iters) {
    __shared__ uint32_t win[NPLANE][PW * PW];
    for (int i = threadIdx.x; i < NPLANE * PW * PW; i += blockDim.x) (&win[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t acc = 0;
    float facc = 0.f;
    for (int it = 0; it < iters; ++it) {
        const uint32_t i = (blockIdx.x * iters + it) * blockDim.x + threadIdx.x;
        const int p = pixel_of<PATTERN>(i, lane);
        const int plane = it % NPLANE;
        uint32_t* w = &win[plane][p];
        const uint32_t v = 1000u + (i & 255u);
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                uint32_t* a = w + dy * PW + dx;
                if (MODE == 0) atomicAdd(a, v + dx);
                else if (MODE == 1) atomicAdd(reinterpret_cast<float*>(a), (float)(v + dx));
                else if (MODE == 2) facc += *reinterpret_cast<volatile float*>(a);
                else *reinterpret_cast<volatile uint32_t*>(a) = v + dx;
            }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NPLANE * PW * PW; i += blockDim.x) acc += (&win[0][0])[i];
    if (acc == 0xdeadbeefu || facc == 1.2345f) out[blockIdx.x] = acc;
}

template <int MODE, int PATTERN>
void run(const char* name, int ctas_per_sm, int iters) {
    uint32_t* out;
    cudaMalloc(&out, 1 << 20);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int grid = sms * ctas_per_sm;
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a);
        k<MODE, PATTERN><<<grid, 256>>>(out, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    const double events = (double)grid * 256 * iters;
    const double lanes = events * 9;
    const double cyc = best * 1e-3 * khz * 1e3;
    printf("%-34s ctas/SM %d  %8.1f us  %7.2f Gevent/s  %6.2f lanes/clk/SM  (%s)\n", name, ctas_per_sm, best * 1e3, events / (best * 1e-3) / 1e9,
           lanes / cyc / sms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    const int iters = 64;
    for (int c : {1, 2, 4, 8}) {
        run<0, 0>("u32 ATOMS.ADD, spread", c, iters);
        run<0, 1>("u32 ATOMS.ADD, lane pairs", c, iters);
        run<0, 2>("u32 ATOMS.ADD, line", c, iters);
    }
    for (int c : {2, 8}) {
        run<1, 0>("f32 CAS loop, spread", c, iters);
        run<1, 2>("f32 CAS loop, line", c, iters);
        run<2, 0>("LDS gather, spread", c, iters);
        run<2, 2>("LDS gather, line", c, iters);
        run<3, 0>("STS, spread", c, iters);
    }
    return 0;
}
