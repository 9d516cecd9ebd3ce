// FP64 / FP32 / shuffle issue rates of one SM sub-partition on this GPU (round 2: explains why float64-heavy kernels crawl).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu && ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int ILP>
__global__ void k_fma(T* out, T a, T b, int iters) {
    T v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = v[i] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename T, int ILP>
double run(int blocks, int threads, int iters) {
    T* out;
    cudaMalloc(&out, sizeof(T) * blocks * threads);
    k_fma<T, ILP><<<blocks, threads>>>(out, (T)1.000001, (T)0.5, iters);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k_fma<T, ILP><<<blocks, threads>>>(out, (T)1.000001, (T)0.5, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    cudaFree(out);
    return 2.0 * blocks * threads * (double)iters * ILP / (ms * 1e-3) / 1e12;   // TFLOP/s
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sm = p.multiProcessorCount;
    printf("%s, %d SMs, %.0f MHz\n", p.name, sm, p.clockRate / 1e3);
    printf("fp64 FMA, 8 independent chains / thread, 1024 thr x 2 CTA / SM : %.2f TFLOP/s\n", run<double, 8>(sm * 2, 1024, 4096));
    printf("fp64 FMA, 1 chain / thread, 1 warp / SMSP (latency)            : %.3f TFLOP/s\n", run<double, 1>(sm, 128, 1 << 16));
    printf("fp32 FMA, 8 independent chains / thread, 1024 thr x 2 CTA / SM : %.2f TFLOP/s\n", run<float, 8>(sm * 2, 1024, 8192));
    printf("fp32 FMA, 1 chain / thread, 1 warp / SMSP (latency)            : %.3f TFLOP/s\n", run<float, 1>(sm, 128, 1 << 16));
    return 0;
}
