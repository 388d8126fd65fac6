// Micro-benchmark: issue throughput of scalar vs packed (f32x2) FP32 instructions on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp32 ubench_fp32.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define UNROLL 8
__device__ __forceinline__ unsigned long long pk(float2 v) { return *reinterpret_cast<unsigned long long *>(&v); }
__device__ __forceinline__ float2 upk(unsigned long long v) { return *reinterpret_cast<float2 *>(&v); }

template <int MODE> __global__ void __launch_bounds__(256) bench(float2 *out, float s) {
    float2 a[UNROLL];
    float2 w = make_float2(s, 1.0f - s);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) a[u] = make_float2(threadIdx.x * 1e-3f + u, blockIdx.x * 1e-3f - u);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (MODE == 0) { a[u].x = a[u].x + w.x; a[u].y = a[u].y + w.y; }                              // 2 FADD
            if (MODE == 1) { unsigned long long r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a[u])), "l"(pk(w))); a[u] = upk(r); }
            if (MODE == 2) { a[u].x = fmaf(a[u].x, w.x, w.y); a[u].y = fmaf(a[u].y, w.y, w.x); }          // 2 FFMA
            if (MODE == 3) { unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(a[u])), "l"(pk(w)), "l"(pk(w))); a[u] = upk(r); }
            if (MODE == 4) { unsigned long long r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a[u])), "l"(pk(w))); a[u] = upk(r); }
        }
    }
    float2 acc = make_float2(0, 0);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { acc.x += a[u].x; acc.y += a[u].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE> void run(const char *name, int flops_per_op) {
    float2 *out;
    int blocks = 148 * 8;
    cudaMalloc(&out, blocks * 256 * sizeof(float2));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE><<<blocks, 256>>>(out, 0.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) bench<MODE><<<blocks, 256>>>(out, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double elem_ops = (double)blocks * 256 * ITERS * UNROLL * 2;          // float lanes updated
    printf("%-10s %8.3f ms  %8.2f G lane-ops/s  %7.2f TFLOP/s\n", name, ms, elem_ops / ms / 1e6, elem_ops * flops_per_op / ms / 1e9);
    cudaFree(out);
}
int main() {
    run<0>("FADD", 1); run<1>("FADD2", 1); run<2>("FFMA", 2); run<3>("FFMA2", 2); run<4>("FMUL2", 1);
    return 0;
}
