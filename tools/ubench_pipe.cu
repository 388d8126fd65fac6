// Ceiling of the pipelined-slab schedule, without any arithmetic: slab k of x is copied into an L2-sized ring slot
// (pass 1), then from the slot to y (pass 2), slabs issued round-robin on NS streams exactly as common.h: PipeScope
// does.  Reports the HBM-level rate (bytes of x + bytes of y) / time, to be compared with a plain device copy.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/ubench_pipe tools/ubench_pipe.cu
//   build/ubench_pipe [total MiB = 16384] [slab MiB = 16] [streams = 3] [inplace = 0]
// inplace = 1: pass 1 writes y's slab and pass 2 rewrites it in place (what the FWHT does).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) copy4(const float4 *__restrict__ in, float4 *__restrict__ out, size_t n) {
    // each thread moves 4 x 16 B, consecutive threads consecutive vectors
    size_t i = (size_t)blockIdx.x * 1024 + threadIdx.x;
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) if (i + 256 * k < n) v[k] = in[i + 256 * k];
#pragma unroll
    for (int k = 0; k < 4; ++k) if (i + 256 * k < n) out[i + 256 * k] = v[k];
}

// merged launch: the first g1 CTAs copy in1 -> out1 (pass 1 of one slab), the others in2 -> out2 (pass 2 of another slab)
__global__ void __launch_bounds__(256) copy4x2(const float4 *__restrict__ in1, float4 *__restrict__ out1, const float4 *__restrict__ in2,
                                               float4 *__restrict__ out2, size_t n, unsigned g1, int interleave) {
    unsigned b = blockIdx.x;
    bool second;
    if (interleave) { second = b & 1; b >>= 1; } else { second = b >= g1; if (second) b -= g1; }
    const float4 *in = second ? in2 : in1;
    float4 *out = second ? out2 : out1;
    if (in == nullptr) return;
    size_t i = (size_t)b * 1024 + threadIdx.x;
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) if (i + 256 * k < n) v[k] = in[i + 256 * k];
#pragma unroll
    for (int k = 0; k < 4; ++k) if (i + 256 * k < n) out[i + 256 * k] = v[k];
}

int main(int argc, char **argv) {
    const size_t total = (size_t)(argc > 1 ? atol(argv[1]) : 16384) << 20, slab = (size_t)(argc > 2 ? atol(argv[2]) : 16) << 20;
    const int ns = argc > 3 ? atoi(argv[3]) : 3, inplace = argc > 4 ? atoi(argv[4]) : 0, merged = argc > 5 ? atoi(argv[5]) : 0;
    float4 *x, *y, *ring;
    CK(cudaMalloc(&x, total)); CK(cudaMalloc(&y, total)); CK(cudaMalloc(&ring, slab * ns * 2));
    CK(cudaMemset(x, 1, total)); CK(cudaMemset(y, 0, total));
    std::vector<cudaStream_t> st(ns);
    for (auto &s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaStream_t main_st; CK(cudaStreamCreate(&main_st));
    cudaEvent_t fork, e0, e1; std::vector<cudaEvent_t> join(ns);
    CK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    for (auto &j : join) CK(cudaEventCreateWithFlags(&j, cudaEventDisableTiming));
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const size_t nv = slab / 16, nslab = total / slab;
    const unsigned grid = (unsigned)((nv + 1023) / 1024);
    auto run = [&]() {
        CK(cudaEventRecord(fork, main_st));
        for (int s = 0; s < ns; ++s) CK(cudaStreamWaitEvent(st[s], fork, 0));
        if (merged) {
            // stream s owns slabs s, s + ns, ...; launch j of a stream = {pass 1 of its slab j, pass 2 of its slab j - 1}
            for (size_t k = 0; k < nslab + ns; ++k) {
                cudaStream_t s = st[k % ns];
                const bool has1 = k < nslab, has2 = k >= (size_t)ns;
                const size_t k2 = k - ns;
                float4 *mid1 = inplace ? y + k * nv : ring + ((k / ns) % 2 * ns + k % ns) * nv;
                float4 *mid2 = inplace ? y + k2 * nv : ring + ((k2 / ns) % 2 * ns + k2 % ns) * nv;
                copy4x2<<<2 * grid, 256, 0, s>>>(has1 ? x + k * nv : nullptr, mid1, has2 ? mid2 : nullptr, has2 ? y + k2 * nv : nullptr, nv, grid, merged == 2);
            }
        } else
        for (size_t k = 0; k < nslab; ++k) {
            cudaStream_t s = st[k % ns];
            float4 *mid = inplace ? y + k * nv : ring + (k % ns) * nv;
            copy4<<<grid, 256, 0, s>>>(x + k * nv, mid, nv);
            copy4<<<grid, 256, 0, s>>>(mid, y + k * nv, nv);
        }
        for (int s = 0; s < ns; ++s) { CK(cudaEventRecord(join[s], st[s])); CK(cudaStreamWaitEvent(main_st, join[s], 0)); }
    };
    run(); run(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0, main_st)); run(); CK(cudaEventRecord(e1, main_st)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    // plain copy x -> y in one launch per 512 MiB for comparison
    float bestc = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0, main_st));
        const size_t chunk = (size_t)512 << 20, cv = chunk / 16;
        for (size_t o = 0; o < total; o += chunk) copy4<<<(unsigned)((cv + 1023) / 1024), 256, 0, main_st>>>(x + o / 16, y + o / 16, cv);
        CK(cudaEventRecord(e1, main_st)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < bestc) bestc = ms;
    }
    printf("total %zu MiB slab %zu MiB streams %d inplace %d merged %d : two-pass pipeline %.3f ms = %.0f GB/s | plain copy %.3f ms = %.0f GB/s | ratio %.3f\n",
           total >> 20, slab >> 20, ns, inplace, merged, best, 2.0 * total / best / 1e6, bestc, 2.0 * total / bestc / 1e6, bestc / best);
    return 0;
}
