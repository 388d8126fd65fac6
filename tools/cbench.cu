// Stand-alone timing harness for the C-ABI (no Python, no torch): dlopen()s a build of libfastmat_b200.so, creates one
// plan, fills a column-major batch on the device and times fmb_plan_apply with CUDA events.  Used for A/B sweeps of kernel
// variants (one short process per configuration; the library reads its FMB_* switches once per process).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -Iinclude -o build/cbench tools/cbench.cu -ldl
//   build/cbench <lib.so> <op> <cols> [reps=5] [seconds=0]
//     op: circ circb fourier fourierb toep toepb kron had blue f16   (b = backward)
//     seconds > 0: additionally run back to back for at least that long and print the sustained figure
// Prints one line: op, cols, ms per apply (best / mean of reps), algorithmic GB/s, fraction of FMB_PEAK_GBS (default
// 6449.7), and two checksums of y (sum |y|^2 and an index-weighted sum) to compare variants with each other.
#include <cuda_profiler_api.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <functional>
#include <string>
#include <vector>

#include "fastmat_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ unsigned hash32(unsigned long long i) {
    unsigned long long z = i + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (unsigned)((z ^ (z >> 31)) >> 32);
}
__global__ void fill_f32(float *p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = ((float)hash32(i) * (1.0f / 4294967296.0f) - 0.5f) * 3.4641016f;      // unit variance
}
__global__ void fill_f64(double *p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = ((double)hash32(i) * (1.0 / 4294967296.0) - 0.5) * 3.4641016151377544;
}
template <typename T> __global__ void checksum(const T *p, size_t n, double *out) {
    double s = 0, w = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double v = (double)p[i];
        s += v * v;
        w += v * (double)((i % 1021) + 1);
    }
    for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); w += __shfl_xor_sync(0xffffffffu, w, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, s); atomicAdd(out + 1, w); }
}

struct Api {
    void *h;
    decltype(&fmb_last_error) last_error;
    decltype(&fmb_fourier_plan_create) fourier;
    decltype(&fmb_circulant_plan_create) circulant;
    decltype(&fmb_toeplitz_plan_create) toeplitz;
    decltype(&fmb_hadamard_plan_create) hadamard;
    decltype(&fmb_kron_fourier_plan_create) kron;
    decltype(&fmb_plan_workspace_bytes) ws_bytes;
    decltype(&fmb_plan_apply) apply;
    decltype(&fmb_plan_destroy) destroy;
    decltype(&fmb_launch_count) launches;
};
template <typename F> static void sym(void *h, const char *n, F &f) {
    f = (F)dlsym(h, n);
    if (!f) { fprintf(stderr, "missing symbol %s\n", n); exit(2); }
}

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s <lib.so> <op> <cols> [reps] [seconds]\n", argv[0]); return 2; }
    const std::string op = argv[2];
    const long cols = atol(argv[3]);
    const int reps = argc > 4 ? atoi(argv[4]) : 5;
    const double seconds = argc > 5 ? atof(argv[5]) : 0.0;
    const double peak = getenv("FMB_PEAK_GBS") ? atof(getenv("FMB_PEAK_GBS")) : 6449.7;
    Api a;
    a.h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!a.h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    sym(a.h, "fmb_last_error", a.last_error); sym(a.h, "fmb_fourier_plan_create", a.fourier);
    sym(a.h, "fmb_circulant_plan_create", a.circulant); sym(a.h, "fmb_toeplitz_plan_create", a.toeplitz);
    sym(a.h, "fmb_hadamard_plan_create", a.hadamard); sym(a.h, "fmb_kron_fourier_plan_create", a.kron);
    sym(a.h, "fmb_plan_workspace_bytes", a.ws_bytes); sym(a.h, "fmb_plan_apply", a.apply);
    sym(a.h, "fmb_plan_destroy", a.destroy); sym(a.h, "fmb_launch_count", a.launches);

    CK(cudaSetDevice(0));
    fmb_plan *plan = nullptr;
    int64_t n_in = 0, n_out = 0;
    int dt = FMB_COMPLEX64, dir = FMB_FORWARD, rc = 0;
    size_t esz = 8;
    double alg_bytes_per_col = 0;
    auto gen = [](size_t n, unsigned seed) {
        std::vector<std::complex<double>> v(n);
        unsigned long long s = seed * 0x9E3779B97F4A7C15ull + 12345;
        for (auto &z : v) {
            s = s * 6364136223846793005ull + 1442695040888963407ull; double re = (double)(s >> 11) / 9007199254740992.0 - 0.5;
            s = s * 6364136223846793005ull + 1442695040888963407ull; double im = (double)(s >> 11) / 9007199254740992.0 - 0.5;
            z = {re, im};
        }
        return v;
    };
    if (op == "circ" || op == "circb") {
        const int64_t N = 1 << 20;
        auto c = gen(N, 1);
        rc = a.circulant(&plan, c.data(), N, 1, 4);
        n_in = n_out = N; dir = op == "circb"; alg_bytes_per_col = 16.0 * N;
    } else if (op == "fourier" || op == "fourierb") {
        const int64_t N = 1 << 20;
        rc = a.fourier(&plan, N, 1, 4);
        n_in = n_out = N; dir = op == "fourierb"; alg_bytes_per_col = 16.0 * N;
    } else if (op == "toep" || op == "toepb") {
        const int64_t n = 1 << 19;
        auto vc = gen(n, 2), vr = gen(n - 1, 3);
        rc = a.toeplitz(&plan, vc.data(), n, vr.data(), n - 1, 1, 4);
        n_in = n_out = n; dir = op == "toepb"; alg_bytes_per_col = 16.0 * n;
    } else if (op == "kron") {
        const int64_t dims[2] = {1024, 1024};
        rc = a.kron(&plan, dims, 2);
        n_in = n_out = 1 << 20; alg_bytes_per_col = 16.0 * (1 << 20);
    } else if (op == "had") {
        rc = a.hadamard(&plan, 20);
        n_in = n_out = 1 << 20; dt = FMB_FLOAT32; esz = 4; alg_bytes_per_col = 8.0 * (1 << 20);
    } else if (op == "blue") {
        rc = a.fourier(&plan, 1000003, 1, 4);
        n_in = n_out = 1000003; alg_bytes_per_col = 16.0 * 1000003;
    } else if (op == "f16") {
        rc = a.fourier(&plan, 1 << 16, 1, 4);
        n_in = n_out = 1 << 16; dt = FMB_COMPLEX128; esz = 16; alg_bytes_per_col = 32.0 * (1 << 16);
    } else { fprintf(stderr, "unknown op %s\n", op.c_str()); return 2; }
    if (rc) { fprintf(stderr, "plan create failed: %s\n", a.last_error()); return 2; }

    const size_t nx = (size_t)n_in * cols, ny = (size_t)n_out * cols;
    void *x, *y, *ws = nullptr;
    CK(cudaMalloc(&x, nx * esz)); CK(cudaMalloc(&y, ny * esz));
    const size_t wx = nx * esz / (dt == FMB_COMPLEX128 ? 8 : 4);
    if (dt == FMB_COMPLEX128) fill_f64<<<1184, 256>>>((double *)x, wx); else fill_f32<<<1184, 256>>>((float *)x, wx);
    CK(cudaMemset(y, 0, ny * esz));
    const int64_t wsb = a.ws_bytes(plan, dir, cols, dt, dt);
    if (wsb < 0) { fprintf(stderr, "workspace query failed: %s\n", a.last_error()); return 2; }
    if (wsb > 0) CK(cudaMalloc(&ws, (size_t)wsb));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    auto run = [&]() {
        int r = a.apply(plan, dir, x, 1, n_in, y, 1, n_out, cols, dt, dt, ws, wsb, st);
        if (r) { fprintf(stderr, "apply failed: %s\n", a.last_error()); exit(2); }
    };
    run(); run(); run();
    CK(cudaStreamSynchronize(st));
    const int64_t l0 = a.launches();
    run();
    const int64_t launches = a.launches() - l0;
    CK(cudaStreamSynchronize(st));
    std::function<void()> direct = run;
    cudaGraphExec_t gexec = nullptr;
    if (getenv("CBENCH_GRAPH") && atoi(getenv("CBENCH_GRAPH"))) {                       // experiment: the whole apply as one CUDA graph launch
        cudaGraph_t graph;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        direct();
        CK(cudaStreamEndCapture(st, &graph));
        CK(cudaGraphInstantiate(&gexec, graph, 0));
        CK(cudaGraphLaunch(gexec, st));
        CK(cudaStreamSynchronize(st));
    }
    auto run2 = [&]() { if (gexec) CK(cudaGraphLaunch(gexec, st)); else direct(); };
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double best = 1e30, sum = 0, enq = 0;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0, st));
        const auto h0 = std::chrono::steady_clock::now();
        run2();
        enq += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, (double)ms); sum += ms;
    }
    double sustained = 0;
    if (seconds > 0) {
        const int n = std::max(1, (int)(seconds * 1e3 / (sum / reps)) + 1);
        CK(cudaEventRecord(e0, st));
        for (int r = 0; r < n; ++r) run2();
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        sustained = ms / n;
    }
    if (getenv("CBENCH_PROFILE")) {                    // one apply inside a profiler range (ncu --replay-mode range: concurrent kernels as they run)
        CK(cudaStreamSynchronize(st));
        CK(cudaProfilerStart());
        run2();
        CK(cudaStreamSynchronize(st));
        CK(cudaProfilerStop());
    }
    double *cs;
    CK(cudaMalloc(&cs, 16)); CK(cudaMemset(cs, 0, 16));
    const size_t wy = ny * esz / (dt == FMB_COMPLEX128 ? 8 : 4);
    if (dt == FMB_COMPLEX128) checksum<double><<<592, 256>>>((const double *)y, wy, cs); else checksum<float><<<592, 256>>>((const float *)y, wy, cs);
    double hcs[2];
    CK(cudaMemcpy(hcs, cs, 16, cudaMemcpyDeviceToHost));
    const double gb = alg_bytes_per_col * cols / 1e9, mean = sum / reps;
    printf("%-8s cols %5ld  best %8.3f ms  mean %8.3f ms  %7.0f GB/s  frac %.3f", op.c_str(), cols, best, mean, gb / mean * 1e3, gb / mean * 1e3 / peak);
    if (seconds > 0) printf("  sustained %8.3f ms frac %.3f", sustained, gb / sustained * 1e3 / peak);
    printf("  enqueue %.3f ms", enq / reps);      // host time inside the apply call (no synchronisation): launch-bound if close to the device time
    printf("  launches %lld  chk %.9e %.9e\n", (long long)launches, hcs[0], hcs[1]);
    if (void (*dump)() = (void (*)())dlsym(a.h, "fmb_debug_dump")) dump();      // instrumented experiment builds only
    a.destroy(plan);
    return 0;
}
