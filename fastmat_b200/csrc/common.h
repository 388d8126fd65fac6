// Host-side helpers shared by the translation units of libfastmat_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>

#include <atomic>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fastmat_b200.h"

namespace fmb {

typedef std::complex<double> cd;

void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;
// internal status (never leaves the library): "this path cannot run here, take the next one"
constexpr int FMB_ERR_FALLBACK = -100;

#define FMB_CUDA_OK(expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            fmb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FMB_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

#define FMB_LAUNCH_OK()                                                                           \
    do {                                                                                          \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess) {                                                                  \
            fmb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FMB_ERR_CUDA;                                                                  \
        }                                                                                         \
        fmb::g_launches.fetch_add(1, std::memory_order_relaxed);                                  \
    } while (0)

// FMB_EMULATE builds (tests/emul only, never the shipped library) run the kernel bodies on host threads so that the
// index logic can be checked in a container without a GPU; "device" memory is then plain host memory.
#ifdef FMB_EMULATE
inline cudaError_t emu_malloc(void **p, size_t n) { *p = malloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t emu_free(void *p) { free(p); return cudaSuccess; }
inline cudaError_t emu_memcpy(void *d, const void *s, size_t n) { memcpy(d, s, n); return cudaSuccess; }
#define FMB_DEV_MALLOC(p, n) fmb::emu_malloc(p, n)
#define FMB_DEV_FREE(p) fmb::emu_free(p)
#define FMB_H2D(d, s, n) fmb::emu_memcpy(d, s, n)
#define FMB_D2H(d, s, n) fmb::emu_memcpy(d, s, n)
void emulate_launch(long long tiles, int nt, size_t smem,
                    const std::function<void(long long, int, int, void *, const std::function<void()> &)> &body);
#else
#define FMB_DEV_MALLOC(p, n) cudaMalloc(p, n)
#define FMB_DEV_FREE(p) cudaFree(p)
#define FMB_H2D(d, s, n) cudaMemcpy(d, s, n, cudaMemcpyHostToDevice)
#define FMB_D2H(d, s, n) cudaMemcpy(d, s, n, cudaMemcpyDeviceToHost)
#endif

struct DevArray {
    void *p = nullptr;
    size_t bytes = 0;
    DevArray() {}
    DevArray(const DevArray &) = delete;
    DevArray &operator=(const DevArray &) = delete;
    ~DevArray() { release(); }
    void release() {
        if (p) FMB_DEV_FREE(p);
        p = nullptr;
        bytes = 0;
    }
    int upload(const void *host, size_t n) {
        release();
        if (n == 0) return FMB_OK;
        FMB_CUDA_OK(FMB_DEV_MALLOC(&p, n));
        bytes = n;
        FMB_CUDA_OK(FMB_H2D(p, host, n));
        return FMB_OK;
    }
    int alloc(size_t n) {
        release();
        if (n == 0) return FMB_OK;
        FMB_CUDA_OK(FMB_DEV_MALLOC(&p, n));
        bytes = n;
        return FMB_OK;
    }
};

inline size_t dtype_size(int dt) {
    switch (dt) {
        case FMB_INT8: return 1;
        case FMB_INT16: return 2;
        case FMB_INT32: return 4;
        case FMB_INT64: return 8;
        case FMB_FLOAT32: return 4;
        case FMB_FLOAT64: return 8;
        case FMB_COMPLEX64: return 8;
        case FMB_COMPLEX128: return 16;
        default: return 0;
    }
}

struct DeviceProps {
    int sm_count = 148;
    int cc_major = 0, cc_minor = 0;
    size_t l2_bytes = 0;
    size_t smem_optin = 0;
    bool ok = false;
};
const DeviceProps &device_props();

// ---- pipelined-slab schedule: an apply whose columns are processed slab by slab (several launches per slab, the
// intermediate of a slab resident in L2) forks the caller's stream into up to FMB_MAX_PIPE internal streams, issues slab
// k on stream(k), and joins them back, so that consecutive slabs overlap and the caller still sees stream order.
constexpr int FMB_MAX_PIPE = 6;
struct PipeScope {
    int ns = 1;
    cudaStream_t caller = nullptr;
    void *pool_ = nullptr;              // the device's stream pool (fft_engine.cu)
    bool open = false;                  // begin() forked the caller's stream and end() has not joined it yet
    int begin(int ns_, cudaStream_t st);
    cudaStream_t stream(int64_t k) const;
    int end();
    ~PipeScope() { if (open) end(); }   // error returns inside a slab loop still join the internal streams
};

// ---- planner (planner.cpp): fastmat/core/cmath.pyx:35-214
int64_t find_optimal_fft_size(int64_t order, int max_stage);
float fft_complexity(int64_t n);
// ---- LFSR sequences (planner.cpp): fastmat/LFSRCirculant.pyx:28-48, :196-222, :277-395
int lfsr_order(uint32_t polynomial);
int64_t lfsr_period(uint32_t polynomial, uint32_t start);
void lfsr_sequences(uint32_t polynomial, uint32_t start, int64_t n, uint32_t *gen_states, uint32_t *tap_states, int8_t *vec_c);

// ---- Plan base
struct PlanBase {
    int kind = 0;
    int64_t num_rows = 0, num_cols = 0;
    virtual ~PlanBase() {}
    virtual int info(fmb_plan_info *out) const = 0;
    virtual int64_t workspace_bytes(int direction, int64_t M, int dt_in, int dt_out) const { return 0; }
    virtual int apply(int direction, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs, int64_t ycs, int64_t M,
                      int dt_in, int dt_out, void *ws, int64_t ws_bytes, cudaStream_t st) const = 0;
};

}  // namespace fmb
