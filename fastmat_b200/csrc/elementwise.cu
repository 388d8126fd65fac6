// Exact element-wise pieces of the apply path: Diag multiply (fastmat/Diag.pyx:149-167 ->
// fastmat/core/cmath.pyx:958-1012 _multiply), the Partial gather / scatter (fastmat/Partial.pyx:268-294),
// complex conjugate (fastmat/core/cmath.pyx:744-840).  All are pure bandwidth: one read, one write.
#include <algorithm>
#include "common.h"
#include "cx.cuh"

namespace fmb {

struct EwParams {
    long long n, M;
    const void *x; long long xrs, xcs;
    void *y; long long yrs, ycs;
    const void *d;           // diag vector / index vector
    int dt_x, dt_d;
    int conj_d;
    int c_fastest;           // 1: consecutive threads walk columns (row-major arrays)
};

// ---- typed load with conversion to the output type ------------------------------------------------------
template <typename O> struct Conv;
template <typename O> FMB_HD O from_real_i(long long v) { return (O)v; }
template <typename O> FMB_HD O from_real_f(double v) { return (O)v; }

template <typename O, bool CPLX> struct Loader {
    // real output types
    static FMB_HD O load(const void *p, int dt, long long i) {
        switch (dt) {
            case FMB_INT8: return (O)((const int8_t *)p)[i];
            case FMB_INT16: return (O)((const int16_t *)p)[i];
            case FMB_INT32: return (O)((const int32_t *)p)[i];
            case FMB_INT64: return (O)((const int64_t *)p)[i];
            case FMB_FLOAT32: return (O)((const float *)p)[i];
            case FMB_FLOAT64: return (O)((const double *)p)[i];
            default: return (O)0;
        }
    }
};
template <typename O> struct Loader<O, true> {
    typedef typename real_of<O>::type S;
    static FMB_HD O load(const void *p, int dt, long long i) {
        switch (dt) {
            case FMB_INT8: return mk<O>((S)((const int8_t *)p)[i], 0);
            case FMB_INT16: return mk<O>((S)((const int16_t *)p)[i], 0);
            case FMB_INT32: return mk<O>((S)((const int32_t *)p)[i], 0);
            case FMB_INT64: return mk<O>((S)((const int64_t *)p)[i], 0);
            case FMB_FLOAT32: return mk<O>((S)((const float *)p)[i], 0);
            case FMB_FLOAT64: return mk<O>((S)((const double *)p)[i], 0);
            case FMB_COMPLEX64: { float2 v = ((const float2 *)p)[i]; return mk<O>((S)v.x, (S)v.y); }
            case FMB_COMPLEX128: { double2 v = ((const double2 *)p)[i]; return mk<O>((S)v.x, (S)v.y); }
            default: return mk<O>(0, 0);
        }
    }
};

template <typename O> struct Mul { static FMB_HD O mul(O a, O b, int) { return a * b; } };
template <> struct Mul<int8_t> { static FMB_HD int8_t mul(int8_t a, int8_t b, int) { return (int8_t)(uint8_t)((unsigned)(uint8_t)a * (unsigned)(uint8_t)b); } };
template <> struct Mul<int16_t> { static FMB_HD int16_t mul(int16_t a, int16_t b, int) { return (int16_t)(uint16_t)((unsigned)(uint16_t)a * (unsigned)(uint16_t)b); } };
template <> struct Mul<int32_t> { static FMB_HD int32_t mul(int32_t a, int32_t b, int) { return (int32_t)((uint32_t)a * (uint32_t)b); } };
template <> struct Mul<int64_t> { static FMB_HD int64_t mul(int64_t a, int64_t b, int) { return (int64_t)((uint64_t)a * (uint64_t)b); } };
template <> struct Mul<float2> { static FMB_HD float2 mul(float2 a, float2 b, int cj) { return cj ? cmulc(a, b) : cmul(a, b); } };
template <> struct Mul<double2> { static FMB_HD double2 mul(double2 a, double2 b, int cj) { return cj ? cmulc(a, b) : cmul(a, b); } };

FMB_HD void ew_decode(const EwParams &p, long long e, long long &n, long long &c) {
    if (p.c_fastest) { n = e / p.M; c = e - n * p.M; }
    else { c = e / p.n; n = e - c * p.n; }
}

template <typename O, bool CPLX> __global__ void __launch_bounds__(256) diag_kernel(const __grid_constant__ EwParams p) {
    const long long total = p.n * p.M;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long n, c;
        ew_decode(p, e, n, c);
        O xv = Loader<O, CPLX>::load(p.x, p.dt_x, n * p.xrs + c * p.xcs);
        O dv = Loader<O, CPLX>::load(p.d, p.dt_d, n);
        ((O *)p.y)[n * p.yrs + c * p.ycs] = Mul<O>::mul(xv, dv, p.conj_d);
    }
}

template <typename B> __global__ void __launch_bounds__(256) gather_kernel(const __grid_constant__ EwParams p) {
    // y[r, c] = x[idx[r], c], r < n
    const long long total = p.n * p.M;
    const long long *idx = (const long long *)p.d;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long r, c;
        ew_decode(p, e, r, c);
        ((B *)p.y)[r * p.yrs + c * p.ycs] = ((const B *)p.x)[__ldg(idx + r) * p.xrs + c * p.xcs];
    }
}

template <typename B> __global__ void __launch_bounds__(256) gather_masked_kernel(const __grid_constant__ EwParams p) {
    // y[r, c] = idx[r] >= 0 ? x[idx[r], c] : 0, r < n   (a scatter expressed through the inverse index)
    const long long total = p.n * p.M;
    const long long *idx = (const long long *)p.d;
    B z;
    memset(&z, 0, sizeof(B));
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long r, c;
        ew_decode(p, e, r, c);
        const long long i = __ldg(idx + r);
        ((B *)p.y)[r * p.yrs + c * p.ycs] = i >= 0 ? ((const B *)p.x)[i * p.xrs + c * p.xcs] : z;
    }
}

template <typename B> __global__ void __launch_bounds__(256) scatter_kernel(const __grid_constant__ EwParams p) {
    // y[idx[r], c] = x[r, c], r < n   (y zero-filled beforehand)
    const long long total = p.n * p.M;
    const long long *idx = (const long long *)p.d;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long r, c;
        ew_decode(p, e, r, c);
        const long long target = __ldg(idx + r);
        // a negative target marks a source row that a LATER source row overwrites (repeated index: the last occurrence
        // wins, as in numpy's y[idx] = x; the plan removes the earlier ones on the host so that nothing races here)
        if (target >= 0) ((B *)p.y)[target * p.yrs + c * p.ycs] = ((const B *)p.x)[r * p.xrs + c * p.xcs];
    }
}

template <typename B> __global__ void __launch_bounds__(256) zero_kernel(const __grid_constant__ EwParams p) {
    const long long total = p.n * p.M;
    B z;
    memset(&z, 0, sizeof(B));
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long r, c;
        ew_decode(p, e, r, c);
        ((B *)p.y)[r * p.yrs + c * p.ycs] = z;
    }
}

template <typename C> __global__ void __launch_bounds__(256) conj_kernel(const __grid_constant__ EwParams p) {
    const long long total = p.n * p.M;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long r, c;
        ew_decode(p, e, r, c);
        C v = ((const C *)p.x)[r * p.xrs + c * p.xcs];
        v.y = -v.y;
        ((C *)p.y)[r * p.yrs + c * p.ycs] = v;
    }
}

static unsigned ew_grid(long long total) {
    long long blocks = (total + 255) / 256;
    long long cap = (long long)device_props().sm_count * 16;        // grid-stride: a multiple of the SM count
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

static EwParams ew_params(int64_t n, int64_t M, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs, int64_t ycs) {
    EwParams p;
    memset(&p, 0, sizeof(p));
    p.n = n; p.M = M; p.x = x; p.xrs = xrs; p.xcs = xcs; p.y = y; p.yrs = yrs; p.ycs = ycs;
    p.c_fastest = (ycs == 1 && M > 1) ? 1 : 0;
    return p;
}

#ifdef FMB_EMULATE
#define FMB_EW_LAUNCH(kernel, p, st) do { set_error("element-wise kernels are not emulated"); return FMB_ERR_NOTIMPL; } while (0)
#else
#define FMB_EW_LAUNCH(kernel, p, st)                                           \
    do {                                                                       \
        if ((p).n * (p).M > 0) {                                               \
            kernel<<<ew_grid((p).n * (p).M), 256, 0, st>>>(p);                 \
            FMB_LAUNCH_OK();                                                   \
        }                                                                      \
    } while (0)
#endif

int diag_apply(const void *d_dev, int dt_d, int64_t n, int conj_d, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs,
               int64_t ycs, int64_t M, int dt_x, int dt_out, cudaStream_t st) {
    EwParams p = ew_params(n, M, x, xrs, xcs, y, yrs, ycs);
    p.d = d_dev; p.dt_x = dt_x; p.dt_d = dt_d; p.conj_d = conj_d;
    const bool in_cplx = dt_x >= FMB_COMPLEX64 || dt_d >= FMB_COMPLEX64;
    if (in_cplx && dt_out < FMB_COMPLEX64) { set_error("Diag: complex operands need a complex output"); return FMB_ERR_TYPE; }
    switch (dt_out) {
        case FMB_INT8: FMB_EW_LAUNCH((diag_kernel<int8_t, false>), p, st); break;
        case FMB_INT16: FMB_EW_LAUNCH((diag_kernel<int16_t, false>), p, st); break;
        case FMB_INT32: FMB_EW_LAUNCH((diag_kernel<int32_t, false>), p, st); break;
        case FMB_INT64: FMB_EW_LAUNCH((diag_kernel<int64_t, false>), p, st); break;
        case FMB_FLOAT32: FMB_EW_LAUNCH((diag_kernel<float, false>), p, st); break;
        case FMB_FLOAT64: FMB_EW_LAUNCH((diag_kernel<double, false>), p, st); break;
        case FMB_COMPLEX64: FMB_EW_LAUNCH((diag_kernel<float2, true>), p, st); break;
        case FMB_COMPLEX128: FMB_EW_LAUNCH((diag_kernel<double2, true>), p, st); break;
        default: set_error("Diag: unsupported output dtype %d", dt_out); return FMB_ERR_TYPE;
    }
    return FMB_OK;
}

template <template <typename> class K> struct BySize {};

// Column-major operands (consecutive threads walk rows): one thread owns one selected row and carries it through CB
// columns, so the index is read once per CB elements, the CB random accesses are independent (all in flight at once)
// and there is no 64-bit division per element.  Blocks are ordered rows-first: the launch sweeps a group of CB columns
// from top to bottom before it moves on, which keeps the randomly addressed side of the copy (CB columns) resident in
// L2 while the other side streams.
template <typename B, bool SCATTER, int CB> __global__ void __launch_bounds__(256) permute_cols_kernel(const __grid_constant__ EwParams p) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.n) return;
    const long long i = __ldg((const long long *)p.d + r);
    if (SCATTER && i < 0) return;                     // a source row that a later one overwrites (repeated index, last wins)
    const bool hole = !SCATTER && i < 0;              // inverse index of a scatter: rows nothing is written to are zero
    const B *src = (const B *)p.x + (SCATTER ? r : (hole ? 0 : i)) * p.xrs;
    B *dst = (B *)p.y + (SCATTER ? i : r) * p.yrs;
    for (long long c0 = (long long)blockIdx.y * CB; c0 < p.M; c0 += (long long)gridDim.y * CB) {
        B v[CB];
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            if (hole) memset(&v[c], 0, sizeof(B));
            else if (c0 + c < p.M) v[c] = src[(c0 + c) * p.xcs];
        }
#pragma unroll
        for (int c = 0; c < CB; ++c)
            if (c0 + c < p.M) dst[(c0 + c) * p.ycs] = v[c];
    }
}

template <typename B, bool SCATTER> static int permute_cols_launch(const EwParams &p, cudaStream_t st) {
#ifdef FMB_EMULATE
    set_error("element-wise kernels are not emulated");
    return FMB_ERR_NOTIMPL;
#else
    constexpr int CB = 4;
    const long long groups = (p.M + CB - 1) / CB;
    dim3 grid((unsigned)((p.n + 255) / 256), (unsigned)(groups < 65535 ? groups : 65535));
    permute_cols_kernel<B, SCATTER, CB><<<grid, 256, 0, st>>>(p);
    FMB_LAUNCH_OK();
    return FMB_OK;
#endif
}

// FMB_EW_TILED=0 keeps the flat grid-stride kernels for column-major operands too (A/B runs)
static bool ew_tiled(const EwParams &p) {
    static const long on = getenv("FMB_EW_TILED") ? atol(getenv("FMB_EW_TILED")) : 1;
    return on && !p.c_fastest && p.n * p.M > 0 && (p.n + 255) / 256 < 0x7fffffffLL;
}

int gather_apply(const void *idx_dev, int64_t nsel, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs, int64_t ycs,
                 int64_t M, int dtype, cudaStream_t st) {
    EwParams p = ew_params(nsel, M, x, xrs, xcs, y, yrs, ycs);
    p.d = idx_dev;
    if (ew_tiled(p)) {
        switch (dtype_size(dtype)) {
            case 1: return permute_cols_launch<uint8_t, false>(p, st);
            case 2: return permute_cols_launch<uint16_t, false>(p, st);
            case 4: return permute_cols_launch<uint32_t, false>(p, st);
            case 8: return permute_cols_launch<uint64_t, false>(p, st);
            case 16: return permute_cols_launch<double2, false>(p, st);
            default: set_error("Partial: unsupported dtype %d", dtype); return FMB_ERR_TYPE;
        }
    }
    switch (dtype_size(dtype)) {
        case 1: FMB_EW_LAUNCH(gather_kernel<uint8_t>, p, st); break;
        case 2: FMB_EW_LAUNCH(gather_kernel<uint16_t>, p, st); break;
        case 4: FMB_EW_LAUNCH(gather_kernel<uint32_t>, p, st); break;
        case 8: FMB_EW_LAUNCH(gather_kernel<uint64_t>, p, st); break;
        case 16: FMB_EW_LAUNCH(gather_kernel<double2>, p, st); break;
        default: set_error("Partial: unsupported dtype %d", dtype); return FMB_ERR_TYPE;
    }
    return FMB_OK;
}

// Scatter through the inverse index (inv[i] = the row of x that lands in row i of y, -1 for none): coalesced stores and
// no separate zero-fill pass; the random side of the copy is a load (measured on B200, LFSRCirculant order 20, 1024
// float32 columns: zero + scatter 9.1 ms, this 4.9 ms).
int scatter_inverse_apply(const void *inv_dev, int64_t ntotal, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs,
                          int64_t ycs, int64_t M, int dtype, cudaStream_t st) {
    EwParams p = ew_params(ntotal, M, x, xrs, xcs, y, yrs, ycs);
    p.d = inv_dev;
    if (ew_tiled(p)) {
        switch (dtype_size(dtype)) {
            case 1: return permute_cols_launch<uint8_t, false>(p, st);
            case 2: return permute_cols_launch<uint16_t, false>(p, st);
            case 4: return permute_cols_launch<uint32_t, false>(p, st);
            case 8: return permute_cols_launch<uint64_t, false>(p, st);
            case 16: return permute_cols_launch<double2, false>(p, st);
            default: set_error("Partial: unsupported dtype %d", dtype); return FMB_ERR_TYPE;
        }
    }
    switch (dtype_size(dtype)) {
        case 1: FMB_EW_LAUNCH(gather_masked_kernel<uint8_t>, p, st); break;
        case 2: FMB_EW_LAUNCH(gather_masked_kernel<uint16_t>, p, st); break;
        case 4: FMB_EW_LAUNCH(gather_masked_kernel<uint32_t>, p, st); break;
        case 8: FMB_EW_LAUNCH(gather_masked_kernel<uint64_t>, p, st); break;
        case 16: FMB_EW_LAUNCH(gather_masked_kernel<double2>, p, st); break;
        default: set_error("Partial: unsupported dtype %d", dtype); return FMB_ERR_TYPE;
    }
    return FMB_OK;
}

int scatter_apply(const void *idx_dev, int64_t nsel, int64_t ntotal, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs,
                  int64_t ycs, int64_t M, int dtype, cudaStream_t st) {
    EwParams z = ew_params(ntotal, M, nullptr, 0, 0, y, yrs, ycs);
    EwParams p = ew_params(nsel, M, x, xrs, xcs, y, yrs, ycs);
    p.d = idx_dev;
    if (ew_tiled(p)) {
        switch (dtype_size(dtype)) {
            case 1: FMB_EW_LAUNCH(zero_kernel<uint8_t>, z, st); return permute_cols_launch<uint8_t, true>(p, st);
            case 2: FMB_EW_LAUNCH(zero_kernel<uint16_t>, z, st); return permute_cols_launch<uint16_t, true>(p, st);
            case 4: FMB_EW_LAUNCH(zero_kernel<uint32_t>, z, st); return permute_cols_launch<uint32_t, true>(p, st);
            case 8: FMB_EW_LAUNCH(zero_kernel<uint64_t>, z, st); return permute_cols_launch<uint64_t, true>(p, st);
            case 16: FMB_EW_LAUNCH(zero_kernel<double2>, z, st); return permute_cols_launch<double2, true>(p, st);
            default: set_error("Partial: unsupported dtype %d", dtype); return FMB_ERR_TYPE;
        }
    }
    switch (dtype_size(dtype)) {
        case 1: FMB_EW_LAUNCH(zero_kernel<uint8_t>, z, st); FMB_EW_LAUNCH(scatter_kernel<uint8_t>, p, st); break;
        case 2: FMB_EW_LAUNCH(zero_kernel<uint16_t>, z, st); FMB_EW_LAUNCH(scatter_kernel<uint16_t>, p, st); break;
        case 4: FMB_EW_LAUNCH(zero_kernel<uint32_t>, z, st); FMB_EW_LAUNCH(scatter_kernel<uint32_t>, p, st); break;
        case 8: FMB_EW_LAUNCH(zero_kernel<uint64_t>, z, st); FMB_EW_LAUNCH(scatter_kernel<uint64_t>, p, st); break;
        case 16: FMB_EW_LAUNCH(zero_kernel<double2>, z, st); FMB_EW_LAUNCH(scatter_kernel<double2>, p, st); break;
        default: set_error("Partial: unsupported dtype %d", dtype); return FMB_ERR_TYPE;
    }
    return FMB_OK;
}

int conj_apply(const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs, int64_t ycs, int64_t n, int64_t M, int dtype, cudaStream_t st) {
    EwParams p = ew_params(n, M, x, xrs, xcs, y, yrs, ycs);
    if (dtype == FMB_COMPLEX64) { FMB_EW_LAUNCH(conj_kernel<float2>, p, st); return FMB_OK; }
    if (dtype == FMB_COMPLEX128) { FMB_EW_LAUNCH(conj_kernel<double2>, p, st); return FMB_OK; }
    // real types: conj is the identity -> plain strided copy through the gather kernel with the identity index
    set_error("conjugate of a real array is the array itself; the caller should not copy");
    return FMB_ERR_TYPE;
}

// copy / cast (x of dt_in) -> (y of dt_out) through the Diag kernel's converting loader with d == 1 is wasteful;
// a dedicated cast: y = (O) x
template <typename O, bool CPLX> __global__ void __launch_bounds__(256) cast_kernel(const __grid_constant__ EwParams p) {
    const long long total = p.n * p.M;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long n, c;
        ew_decode(p, e, n, c);
        ((O *)p.y)[n * p.yrs + c * p.ycs] = Loader<O, CPLX>::load(p.x, p.dt_x, n * p.xrs + c * p.xcs);
    }
}

int cast_apply(const void *x, int64_t xrs, int64_t xcs, int dt_x, void *y, int64_t yrs, int64_t ycs, int dt_out, int64_t n, int64_t M,
               cudaStream_t st) {
    EwParams p = ew_params(n, M, x, xrs, xcs, y, yrs, ycs);
    p.dt_x = dt_x;
    if (dt_x >= FMB_COMPLEX64 && dt_out < FMB_COMPLEX64) { set_error("cast: complex -> real is not allowed"); return FMB_ERR_TYPE; }
    switch (dt_out) {
        case FMB_INT8: FMB_EW_LAUNCH((cast_kernel<int8_t, false>), p, st); break;
        case FMB_INT16: FMB_EW_LAUNCH((cast_kernel<int16_t, false>), p, st); break;
        case FMB_INT32: FMB_EW_LAUNCH((cast_kernel<int32_t, false>), p, st); break;
        case FMB_INT64: FMB_EW_LAUNCH((cast_kernel<int64_t, false>), p, st); break;
        case FMB_FLOAT32: FMB_EW_LAUNCH((cast_kernel<float, false>), p, st); break;
        case FMB_FLOAT64: FMB_EW_LAUNCH((cast_kernel<double, false>), p, st); break;
        case FMB_COMPLEX64: FMB_EW_LAUNCH((cast_kernel<float2, true>), p, st); break;
        case FMB_COMPLEX128: FMB_EW_LAUNCH((cast_kernel<double2, true>), p, st); break;
        default: set_error("cast: unsupported output dtype %d", dt_out); return FMB_ERR_TYPE;
    }
    return FMB_OK;
}

// ---- fused proximal-gradient update of ISTA / FISTA (fastmat/algorithms/ISTA.py:113-123, :150-158):
//          step = x - numL * grad                      (grad == nullptr: step = x, a plain soft threshold)
//          m    = max(|step| - alpha, 0)
//          xnew = m / (m + alpha) * step
// one read of x and grad, one write of step (optional) and xnew, instead of the reference's seven array sweeps.
template <typename S> struct IstaAbs {
    static __device__ __forceinline__ S of(S v) { return fabs(v); }
    static __device__ __forceinline__ S scale(S v, S f) { return v * f; }
    static __device__ __forceinline__ S axpy(S x, S a, S g) { return x - a * g; }
};
template <> struct IstaAbs<float2> {
    static __device__ __forceinline__ float of(float2 v) { return hypotf(v.x, v.y); }
    static __device__ __forceinline__ float2 scale(float2 v, float f) { return make_float2(v.x * f, v.y * f); }
    static __device__ __forceinline__ float2 axpy(float2 x, float a, float2 g) { return make_float2(x.x - a * g.x, x.y - a * g.y); }
};
template <> struct IstaAbs<double2> {
    static __device__ __forceinline__ double of(double2 v) { return hypot(v.x, v.y); }
    static __device__ __forceinline__ double2 scale(double2 v, double f) { return make_double2(v.x * f, v.y * f); }
    static __device__ __forceinline__ double2 axpy(double2 x, double a, double2 g) { return make_double2(x.x - a * g.x, x.y - a * g.y); }
};
template <typename V, typename S>
__global__ void __launch_bounds__(256) ista_step_kernel(const V *__restrict__ x, const V *__restrict__ grad, V *__restrict__ step_out,
                                                        V *__restrict__ x_out, long long count, S numL, S alpha) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long long)gridDim.x * blockDim.x) {
        V s = x[e];
        if (grad != nullptr) s = IstaAbs<V>::axpy(s, numL, grad[e]);
        if (step_out != nullptr) step_out[e] = s;
        const S m = fmax(IstaAbs<V>::of(s) - alpha, (S)0);
        x_out[e] = IstaAbs<V>::scale(s, m / (m + alpha));
    }
}

int ista_step_apply(const void *x, const void *grad, void *step_out, void *x_out, int64_t count, double numL, double alpha, int dtype,
                    cudaStream_t st) {
#ifdef FMB_EMULATE
    set_error("element-wise kernels are not emulated");
    return FMB_ERR_NOTIMPL;
#else
    if (count <= 0) return FMB_OK;
    const unsigned grid = ew_grid(count);
    switch (dtype) {
        case FMB_FLOAT32:
            ista_step_kernel<float, float><<<grid, 256, 0, st>>>((const float *)x, (const float *)grad, (float *)step_out, (float *)x_out, count, (float)numL, (float)alpha);
            break;
        case FMB_FLOAT64:
            ista_step_kernel<double, double><<<grid, 256, 0, st>>>((const double *)x, (const double *)grad, (double *)step_out, (double *)x_out, count, numL, alpha);
            break;
        case FMB_COMPLEX64:
            ista_step_kernel<float2, float><<<grid, 256, 0, st>>>((const float2 *)x, (const float2 *)grad, (float2 *)step_out, (float2 *)x_out, count, (float)numL, (float)alpha);
            break;
        case FMB_COMPLEX128:
            ista_step_kernel<double2, double><<<grid, 256, 0, st>>>((const double2 *)x, (const double2 *)grad, (double2 *)step_out, (double2 *)x_out, count, numL, alpha);
            break;
        default: set_error("ista_step: unsupported dtype %d (float32/64, complex64/128)", dtype); return FMB_ERR_TYPE;
    }
    FMB_LAUNCH_OK();
    return FMB_OK;
#endif
}

// ---- OMP atom selection (fastmat/algorithms/OMP.pyx:196-199: argmax(abs(C^H r)) per column) as ONE sweep: the magnitude
// array is never materialised.  Column-major input, `chunks` CTAs per column write (value, index) partials, a second tiny
// kernel picks the winner; ties go to the lowest index, like np.argmax.
template <typename V> struct Mag2;
template <> struct Mag2<float> { typedef float S; static __device__ __forceinline__ float of(float v) { return v * v; } };
template <> struct Mag2<double> { typedef double S; static __device__ __forceinline__ double of(double v) { return v * v; } };
template <> struct Mag2<float2> { typedef float S; static __device__ __forceinline__ float of(float2 v) { return v.x * v.x + v.y * v.y; } };
template <> struct Mag2<double2> { typedef double S; static __device__ __forceinline__ double of(double2 v) { return v.x * v.x + v.y * v.y; } };
constexpr int ARGMAX_CHUNKS = 8, ARGMAX_NT = 512;
struct ArgmaxPartial { double val; long long idx; };

template <typename V>
__global__ void __launch_bounds__(ARGMAX_NT) abs_argmax_kernel(const V *__restrict__ x, long long rows, long long col_stride,
                                                               ArgmaxPartial *__restrict__ part) {
    typedef typename Mag2<V>::S S;
    const long long col = blockIdx.y, per = (rows + ARGMAX_CHUNKS - 1) / ARGMAX_CHUNKS;
    const long long r0 = blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
    const V *xc = x + col * col_stride;
    S best = (S)-1;
    long long bi = r0;
    for (long long r = r0 + threadIdx.x; r < r1; r += ARGMAX_NT) {
        const S m = Mag2<V>::of(xc[r]);
        if (m > best) { best = m; bi = r; }               // ascending r within a thread: the first maximum is kept
    }
    __shared__ double sv[ARGMAX_NT];
    __shared__ long long si[ARGMAX_NT];
    sv[threadIdx.x] = (double)best; si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = ARGMAX_NT / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const double v2 = sv[threadIdx.x + o];
            const long long i2 = si[threadIdx.x + o];
            if (v2 > sv[threadIdx.x] || (v2 == sv[threadIdx.x] && i2 < si[threadIdx.x])) { sv[threadIdx.x] = v2; si[threadIdx.x] = i2; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[col * ARGMAX_CHUNKS + blockIdx.x].val = sv[0]; part[col * ARGMAX_CHUNKS + blockIdx.x].idx = si[0]; }
}
__global__ void abs_argmax_final_kernel(const ArgmaxPartial *__restrict__ part, long long cols, long long *__restrict__ out) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    double best = -1.0;
    long long bi = 0;
    for (int k = 0; k < ARGMAX_CHUNKS; ++k) {             // chunks in ascending row order: strict > keeps the first maximum
        const ArgmaxPartial p = part[c * ARGMAX_CHUNKS + k];
        if (p.val > best) { best = p.val; bi = p.idx; }
    }
    out[c] = bi;
}

int64_t abs_argmax_workspace_bytes(int64_t cols) { return cols * ARGMAX_CHUNKS * (int64_t)sizeof(ArgmaxPartial); }

int abs_argmax_apply(const void *x, int64_t rows, int64_t cols, int64_t col_stride, int dtype, int64_t *out, void *ws, int64_t ws_bytes,
                     cudaStream_t st) {
#ifdef FMB_EMULATE
    set_error("element-wise kernels are not emulated");
    return FMB_ERR_NOTIMPL;
#else
    if (rows <= 0 || cols <= 0) return FMB_OK;
    if (ws == nullptr || ws_bytes < abs_argmax_workspace_bytes(cols)) { set_error("abs_argmax: workspace too small"); return FMB_ERR_WORKSPACE; }
    if (cols > 65535) { set_error("abs_argmax: more than 65535 columns per call"); return FMB_ERR_VALUE; }
    const dim3 grid(ARGMAX_CHUNKS, (unsigned)cols);
    ArgmaxPartial *part = (ArgmaxPartial *)ws;
    switch (dtype) {
        case FMB_FLOAT32: abs_argmax_kernel<float><<<grid, ARGMAX_NT, 0, st>>>((const float *)x, rows, col_stride, part); break;
        case FMB_FLOAT64: abs_argmax_kernel<double><<<grid, ARGMAX_NT, 0, st>>>((const double *)x, rows, col_stride, part); break;
        case FMB_COMPLEX64: abs_argmax_kernel<float2><<<grid, ARGMAX_NT, 0, st>>>((const float2 *)x, rows, col_stride, part); break;
        case FMB_COMPLEX128: abs_argmax_kernel<double2><<<grid, ARGMAX_NT, 0, st>>>((const double2 *)x, rows, col_stride, part); break;
        default: set_error("abs_argmax: unsupported dtype %d (float32/64, complex64/128)", dtype); return FMB_ERR_TYPE;
    }
    FMB_LAUNCH_OK();
    abs_argmax_final_kernel<<<(unsigned)((cols + 127) / 128), 128, 0, st>>>(part, cols, (long long *)out);
    FMB_LAUNCH_OK();
    return FMB_OK;
#endif
}

// ---- batched Gram-Schmidt step of OMP's incremental QR (fastmat_b200/algorithms/OMP.py): `batches` independent problems,
// problem l holds k orthonormal rows Q[l, j, :] (j < k) of length n and a new vector v[l, :]:
//      project :  coef[l, j] = sum_n conj(Q[l, j, n]) * v[l, n]
//      subtract:  v[l, n]   -= sum_j coef[l, j] * Q[l, j, n]
// Two memory-bound sweeps over the used part of Q per orthogonalisation (cuBLAS batched products of these skinny shapes
// in complex128 took ~40 ms per call at k ~ 16, n = 2^16, 128 batches; these take the time of reading Q).
template <typename V> struct GsOps;
template <> struct GsOps<float> {
    static __device__ __forceinline__ float cmac(float acc, float q, float v) { return acc + q * v; }       // conj(q) * v
    static __device__ __forceinline__ float mac(float acc, float c, float q) { return acc + c * q; }
    static __device__ __forceinline__ float zero() { return 0.f; }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
};
template <> struct GsOps<double> {
    static __device__ __forceinline__ double cmac(double acc, double q, double v) { return acc + q * v; }
    static __device__ __forceinline__ double mac(double acc, double c, double q) { return acc + c * q; }
    static __device__ __forceinline__ double zero() { return 0.0; }
    static __device__ __forceinline__ double add(double a, double b) { return a + b; }
    static __device__ __forceinline__ double sub(double a, double b) { return a - b; }
};
template <> struct GsOps<float2> {
    static __device__ __forceinline__ float2 cmac(float2 a, float2 q, float2 v) { return make_float2(a.x + q.x * v.x + q.y * v.y, a.y + q.x * v.y - q.y * v.x); }
    static __device__ __forceinline__ float2 mac(float2 a, float2 c, float2 q) { return make_float2(a.x + c.x * q.x - c.y * q.y, a.y + c.x * q.y + c.y * q.x); }
    static __device__ __forceinline__ float2 zero() { return make_float2(0.f, 0.f); }
    static __device__ __forceinline__ float2 add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
    static __device__ __forceinline__ float2 sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
};
template <> struct GsOps<double2> {
    static __device__ __forceinline__ double2 cmac(double2 a, double2 q, double2 v) { return make_double2(a.x + q.x * v.x + q.y * v.y, a.y + q.x * v.y - q.y * v.x); }
    static __device__ __forceinline__ double2 mac(double2 a, double2 c, double2 q) { return make_double2(a.x + c.x * q.x - c.y * q.y, a.y + c.x * q.y + c.y * q.x); }
    static __device__ __forceinline__ double2 zero() { return make_double2(0.0, 0.0); }
    static __device__ __forceinline__ double2 add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
    static __device__ __forceinline__ double2 sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
};
constexpr int GS_NT = 256, GS_MAXK = 256;

template <typename V>
__global__ void __launch_bounds__(GS_NT) gs_project_kernel(const V *__restrict__ q, long long q_bs, long long q_rs, const V *__restrict__ v,
                                                           long long v_bs, long long n, V *__restrict__ coef, long long coef_bs) {
    const long long l = blockIdx.y;
    const int j = blockIdx.x;
    const V *qr = q + l * q_bs + j * q_rs, *vr = v + l * v_bs;
    V acc = GsOps<V>::zero();
    for (long long e = threadIdx.x; e < n; e += GS_NT) acc = GsOps<V>::cmac(acc, qr[e], vr[e]);
    __shared__ V red[GS_NT];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = GS_NT / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] = GsOps<V>::add(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) coef[l * coef_bs + j] = red[0];
}
template <typename V>
__global__ void __launch_bounds__(GS_NT) gs_subtract_kernel(const V *__restrict__ q, long long q_bs, long long q_rs, V *__restrict__ v,
                                                            long long v_bs, long long n, const V *__restrict__ coef, long long coef_bs, int k) {
    const long long l = blockIdx.y;
    __shared__ V c[GS_MAXK];
    for (int j = threadIdx.x; j < k; j += GS_NT) c[j] = coef[l * coef_bs + j];
    __syncthreads();
    const V *qb = q + l * q_bs;
    V *vr = v + l * v_bs;
    for (long long e = (long long)blockIdx.x * GS_NT + threadIdx.x; e < n; e += (long long)gridDim.x * GS_NT) {
        V acc = GsOps<V>::zero();
        for (int j = 0; j < k; ++j) acc = GsOps<V>::mac(acc, c[j], qb[j * q_rs + e]);
        vr[e] = GsOps<V>::sub(vr[e], acc);
    }
}

template <typename V>
static int gs_step_t(int subtract, const void *q, int64_t q_bs, int64_t q_rs, int k, void *v, int64_t v_bs, int64_t n, int64_t batches,
                     void *coef, int64_t coef_bs, cudaStream_t st) {
    if (!subtract) {
        gs_project_kernel<V><<<dim3((unsigned)k, (unsigned)batches), GS_NT, 0, st>>>((const V *)q, q_bs, q_rs, (const V *)v, v_bs, n, (V *)coef, coef_bs);
    } else {
        const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + GS_NT - 1) / GS_NT, 64));
        gs_subtract_kernel<V><<<dim3(gx, (unsigned)batches), GS_NT, 0, st>>>((const V *)q, q_bs, q_rs, (V *)v, v_bs, n, (const V *)coef, coef_bs, k);
    }
    FMB_LAUNCH_OK();
    return FMB_OK;
}

int gs_step_apply(int subtract, const void *q, int64_t q_bs, int64_t q_rs, int k, void *v, int64_t v_bs, int64_t n, int64_t batches,
                  void *coef, int64_t coef_bs, int dtype, cudaStream_t st) {
#ifdef FMB_EMULATE
    set_error("element-wise kernels are not emulated");
    return FMB_ERR_NOTIMPL;
#else
    if (k <= 0 || n <= 0 || batches <= 0) return FMB_OK;
    if (k > GS_MAXK || batches > 65535) { set_error("gs_step: at most %d basis vectors and 65535 problems per call", GS_MAXK); return FMB_ERR_VALUE; }
    switch (dtype) {
        case FMB_FLOAT32: return gs_step_t<float>(subtract, q, q_bs, q_rs, k, v, v_bs, n, batches, coef, coef_bs, st);
        case FMB_FLOAT64: return gs_step_t<double>(subtract, q, q_bs, q_rs, k, v, v_bs, n, batches, coef, coef_bs, st);
        case FMB_COMPLEX64: return gs_step_t<float2>(subtract, q, q_bs, q_rs, k, v, v_bs, n, batches, coef, coef_bs, st);
        case FMB_COMPLEX128: return gs_step_t<double2>(subtract, q, q_bs, q_rs, k, v, v_bs, n, batches, coef, coef_bs, st);
        default: set_error("gs_step: unsupported dtype %d (float32/64, complex64/128)", dtype); return FMB_ERR_TYPE;
    }
#endif
}

// ---- out[j * ldb + i] = in[i * lda + j], i < ni, j < nj (32 x 32 tiles through shared memory, both sides coalesced).
// The FFT engine uses it to bring a chunk of a row-major (batch-contiguous, torch default) operand into the column-major
// layout its fast paths work on, and the result back (fft_engine.cu: run_t).
template <typename V>
__global__ void __launch_bounds__(256) transpose_kernel(const V *__restrict__ in, long long lda, V *__restrict__ out, long long ldb,
                                                        long long ni, long long nj) {
    __shared__ V tile[32][33];
    const long long i0 = (long long)blockIdx.y * 32, j0 = (long long)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = ty; r < 32; r += 8)
        if (i0 + r < ni && j0 + tx < nj) tile[r][tx] = in[(i0 + r) * lda + j0 + tx];
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8)
        if (j0 + r < nj && i0 + tx < ni) out[(j0 + r) * ldb + i0 + tx] = tile[tx][r];
}

int transpose_apply(const void *in, int64_t lda, void *out, int64_t ldb, int64_t ni, int64_t nj, size_t esize, cudaStream_t st) {
#ifdef FMB_EMULATE
    set_error("element-wise kernels are not emulated");
    return FMB_ERR_NOTIMPL;
#else
    if (ni <= 0 || nj <= 0) return FMB_OK;
    const long long gy = (ni + 31) / 32, gx = (nj + 31) / 32;
    if (gy > 65535) {
        // rows are the long dimension: put them on grid.x by swapping the roles of the block indices is not possible with a
        // fixed kernel, so walk the rows in slices of 65535 tiles
        for (long long t0 = 0; t0 < gy; t0 += 65535) {
            const long long nt = std::min<long long>(65535, gy - t0), r0 = t0 * 32, nr = std::min<long long>(nt * 32, ni - r0);
            const dim3 grid((unsigned)gx, (unsigned)nt);
            if (esize == 8) transpose_kernel<float2><<<grid, 256, 0, st>>>((const float2 *)in + r0 * lda, lda, (float2 *)out + r0, ldb, nr, nj);
            else transpose_kernel<double2><<<grid, 256, 0, st>>>((const double2 *)in + r0 * lda, lda, (double2 *)out + r0, ldb, nr, nj);
            FMB_LAUNCH_OK();
        }
        return FMB_OK;
    }
    const dim3 grid((unsigned)gx, (unsigned)gy);
    if (esize == 8) transpose_kernel<float2><<<grid, 256, 0, st>>>((const float2 *)in, lda, (float2 *)out, ldb, ni, nj);
    else if (esize == 16) transpose_kernel<double2><<<grid, 256, 0, st>>>((const double2 *)in, lda, (double2 *)out, ldb, ni, nj);
    else { set_error("transpose: element size %d", (int)esize); return FMB_ERR_TYPE; }
    FMB_LAUNCH_OK();
    return FMB_OK;
#endif
}

}  // namespace fmb
