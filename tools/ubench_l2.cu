// Micro-benchmark: L2 / HBM bandwidth seen by plain SM loads and stores at several footprints (is the three-pass
// convolution - 48 GB of L2-level traffic per 1024 columns - near an L2 limit?), and the cost of 64-byte segments.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/ubench_l2 tools/ubench_l2.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

// every thread reads 16-byte words, grid-stride over `n4` words, `rounds` times
__global__ void __launch_bounds__(512) rd(const float4 *p, size_t n4, int rounds, float *sink) {
    float acc = 0;
    for (int r = 0; r < rounds; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            float4 v;
            asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i));
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 1.2345e30f) *sink = acc;
}
__global__ void __launch_bounds__(512) cp(const float4 *p, float4 *q, size_t n4, int rounds) {
    for (int r = 0; r < rounds; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            float4 v;
            asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i));
            q[i] = v;
        }
}
// 8-byte loads where a warp covers `seg` contiguous bytes per row and 256/seg rows `stride` bytes apart (seg = 64: the
// access shape of the strided FFT passes; seg = 256: fully contiguous)
__global__ void __launch_bounds__(256) rd_seg(const float2 *p, size_t n2, int seg, size_t stride2, int rounds, float *sink) {
    float acc = 0;
    const int lanes = seg / 8, lane = threadIdx.x & 31, sub = lane % lanes, row = lane / lanes, rows = 32 / lanes;
    const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5), w = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    // the buffer is a matrix of rows of `stride2` elements; a warp-iteration reads a (rows x lanes) patch
    const size_t patches_per_rowgroup = stride2 / lanes, rowgroups = n2 / (stride2 * rows), total = patches_per_rowgroup * rowgroups;
    for (int r = 0; r < rounds; ++r)
        for (size_t t = w; t < total; t += warps) {
            const size_t rg = t / patches_per_rowgroup, pc = t % patches_per_rowgroup;
            const float2 *q = p + (rg * rows + row) * stride2 + pc * lanes + sub;
            float2 v;
            asm volatile("ld.global.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(q));
            acc += v.x + v.y;
        }
    if (acc == 1.2345e30f) *sink = acc;
}
int main() {
    const size_t maxb = (size_t)2 << 30;
    float4 *a, *b; float *sink;
    CK(cudaMalloc(&a, maxb)); CK(cudaMalloc(&b, maxb)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(a, 0, maxb)); CK(cudaMemset(b, 0, maxb));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const size_t sizes[] = {(size_t)8 << 20, (size_t)16 << 20, (size_t)32 << 20, (size_t)48 << 20, (size_t)64 << 20, (size_t)96 << 20, (size_t)256 << 20, (size_t)2 << 30};
    for (size_t s : sizes) {
        const int rounds = (int)(((size_t)8 << 30) / s); float ms;
        rd<<<148 * 4, 512>>>(a, s / 16, 2, sink); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); rd<<<148 * 4, 512>>>(a, s / 16, rounds, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("read  footprint %5zu MiB: %8.1f GB/s\n", s >> 20, (double)s * rounds / ms / 1e6);
        cp<<<148 * 4, 512>>>(a, b, s / 32, 2); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); cp<<<148 * 4, 512>>>(a, b, s / 32, rounds); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("copy  footprint %5zu MiB (half read, half written): %8.1f GB/s (read+write)\n", s >> 20, (double)s * rounds / ms / 1e6);
    }
    for (size_t s : {(size_t)32 << 20, (size_t)2 << 30})
        for (int seg : {32, 64, 128, 256}) {
            const int rounds = (int)(((size_t)4 << 30) / s); float ms;
            rd_seg<<<148 * 8, 256>>>((const float2 *)a, s / 8, seg, 1024, 1, sink); CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0)); rd_seg<<<148 * 8, 256>>>((const float2 *)a, s / 8, seg, 1024, rounds, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("read  footprint %5zu MiB in %3d-byte segments, 8 KiB apart: %8.1f GB/s\n", s >> 20, seg, (double)s * rounds / ms / 1e6);
        }
    return 0;
}
