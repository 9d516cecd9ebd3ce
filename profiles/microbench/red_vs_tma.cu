// Microbenchmark: cost of adding one 48-byte (12 x f32) record to a pseudo-random record in an L2-resident array on B200:
//   (a) 2 x red.global.add.v4.f32 + 1 x red.global.add.f32   (what k_splat9 does)
//   (b) 3 x red.global.add.v4.f32
//   (c) 9 x red.global.add.f32 (scalar, same record)
//   (d) 1 x cp.reduce.async.bulk.global.shared::cta.add.f32 of 48 bytes (TMA reduction), per thread
//   (e) 9 x red.global.add.f64 to 3 rows (the float64 reference layout)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_vs_tma red_vs_tma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// locality model: consecutive threads hit records within a small neighbourhood (like pixel-sorted events warped by a flow)
__device__ __forceinline__ uint32_t rec_index(uint32_t i, uint32_t n_rec, int W) {
    uint32_t base = (uint32_t)(((uint64_t)(i >> 5) * 2654435761ull) % n_rec);   // per-warp anchor
    uint32_t h = hash32(i);
    int dx = (int)(h & 31) - 16, dy = (int)((h >> 5) & 15) - 8;
    int64_t r = (int64_t)base + dy * W + dx;
    if (r < 0) r += n_rec; if (r >= n_rec) r -= n_rec;
    return (uint32_t)r;
}

template <int STRIDE, int BYTES, int DEPTH>
__global__ void __launch_bounds__(128) k_tma(float* C, uint32_t n_rec, int W, uint32_t n_ops) {
    __shared__ __align__(128) float slot[DEPTH][128][16];
    int d = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ops; i += gridDim.x * blockDim.x) {
        const uint32_t r = rec_index(i, n_rec, W);
        float v = 1.0f + (float)(i & 7);
        float* rec = C + (size_t)r * STRIDE;
        if (d == DEPTH) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); d = 0; }
#pragma unroll
        for (int k2 = 0; k2 < BYTES / 4; ++k2) slot[d][threadIdx.x][k2] = v;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        uint32_t sa = (uint32_t)__cvta_generic_to_shared(&slot[d][threadIdx.x][0]);
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(rec), "r"(sa), "n"(BYTES) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        ++d;
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <int STRIDE, int BYTES, int DEPTH>
float run_tma(float* C, uint32_t n_rec, int W, uint32_t n_ops, int grid, const char* name) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 6; ++it) {
        cudaMemsetAsync(C, 0, (size_t)n_rec * 64);
        cudaEventRecord(a);
        k_tma<STRIDE, BYTES, DEPTH><<<grid * 2, 128>>>(C, n_rec, W, n_ops);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-52s grid %5d %8.1f us  %6.2f Gops/s  (%s)\n", name, grid, best * 1e3, n_ops / (best * 1e-3) / 1e9, cudaGetErrorString(e));
    return best;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* C, double* D, uint32_t n_rec, int W, uint32_t n_ops) {
    __shared__ __align__(16) float slot[256][12];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ops; i += gridDim.x * blockDim.x) {
        const uint32_t r = rec_index(i, n_rec, W);
        float v = 1.0f + (float)(i & 7);
        float* rec = C + (size_t)r * 12;
        if (MODE == 0) {
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(rec), "f"(v) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(rec + 4), "f"(v) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(rec + 8), "f"(v) : "memory");
        } else if (MODE == 1) {
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(rec), "f"(v) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(rec + 4), "f"(v) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(rec + 8), "f"(v) : "memory");
        } else if (MODE == 2) {
#pragma unroll
            for (int k2 = 0; k2 < 9; ++k2) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(rec + k2), "f"(v) : "memory");
        } else if (MODE == 3) {
#pragma unroll
            for (int k2 = 0; k2 < 12; ++k2) slot[threadIdx.x][k2] = v;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            uint32_t sa = (uint32_t)__cvta_generic_to_shared(&slot[threadIdx.x][0]);
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 48;" ::"l"(rec), "r"(sa) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        } else if (MODE == 4) {
            // 9 scalar f64 reductions into a plain [H][W] f64 image: 3 rows x 3 columns around pixel r
#pragma unroll
            for (int j = -1; j <= 1; ++j)
#pragma unroll
                for (int i2 = -1; i2 <= 1; ++i2) {
                    int64_t q = (int64_t)r + j * W + i2;
                    if (q < 0) q += n_rec; if (q >= n_rec) q -= n_rec;
                    atomicAdd(D + q, (double)v);
                }
        } else if (MODE == 5) {   // 2 x v4 only (8 moments), to see the per-op scaling
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(rec), "f"(v) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(rec + 4), "f"(v) : "memory");
        } else if (MODE == 6) {   // 1 x v4
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(rec), "f"(v) : "memory");
        }
    }
}

template <int MODE>
float run(float* C, double* D, uint32_t n_rec, int W, uint32_t n_ops, const char* name, int grid = 148 * 8) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 6; ++it) {
        cudaMemsetAsync(C, 0, (size_t)n_rec * 48); cudaMemsetAsync(D, 0, (size_t)n_rec * 8);
        cudaEventRecord(a);
        k<MODE><<<grid, 256>>>(C, D, n_rec, W, n_ops);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-44s %8.1f us  %6.2f Gops/s  (%s)\n", name, best * 1e3, n_ops / (best * 1e-3) / 1e9, cudaGetErrorString(e));
    return best;
}

int main() {
    const int W = 640, H = 480, R = 3;
    const uint32_t n_rec = W * H * R, n_ops = 6000000;
    float* C; double* D;
    cudaMalloc(&C, (size_t)n_rec * 64); cudaMalloc(&D, (size_t)n_rec * 8);
    run<0>(C, D, n_rec, W, n_ops, "2x RED.v4.f32 + 1x RED.f32 (k_splat9)");
    run<1>(C, D, n_rec, W, n_ops, "3x RED.v4.f32");
    run<5>(C, D, n_rec, W, n_ops, "2x RED.v4.f32");
    run<6>(C, D, n_rec, W, n_ops, "1x RED.v4.f32");
    run<2>(C, D, n_rec, W, n_ops, "9x RED.f32 same record");
    run<4>(C, D, n_rec, W, n_ops, "9x RED.f64 over 3 rows (reference layout)");
    run<3>(C, D, n_rec, W, n_ops, "1x cp.reduce.async.bulk 48 B (TMA)");
    run<1>(C, D, n_rec, W, n_ops, "3x RED.v4.f32, 74 SMs x 8 CTAs", 74 * 8);
    run<1>(C, D, n_rec, W, n_ops, "3x RED.v4.f32, 37 SMs x 8 CTAs", 37 * 8);
    run<6>(C, D, n_rec, W, n_ops, "1x RED.v4.f32, 74 x 8 CTAs", 74 * 8);
    run_tma<12, 48, 1>(C, n_rec, W, n_ops, 148 * 8, "TMA 48 B, stride 48, depth 1");
    run_tma<12, 48, 4>(C, n_rec, W, n_ops, 148 * 4, "TMA 48 B, stride 48, depth 4");
    run_tma<16, 48, 4>(C, n_rec, W, n_ops, 148 * 4, "TMA 48 B, stride 64, depth 4");
    run_tma<16, 64, 4>(C, n_rec, W, n_ops, 148 * 4, "TMA 64 B, stride 64, depth 4");
    run_tma<16, 32, 4>(C, n_rec, W, n_ops, 148 * 4, "TMA 32 B, stride 64, depth 4");
    run_tma<16, 16, 4>(C, n_rec, W, n_ops, 148 * 4, "TMA 16 B, stride 64, depth 4");
    run_tma<16, 64, 4>(C, n_rec, W, n_ops, 74 * 4, "TMA 64 B, stride 64, depth 4, 74x4 CTAs");
    float h[12]; cudaMemcpy(h, C, 48, cudaMemcpyDeviceToHost);
    return 0;
}
