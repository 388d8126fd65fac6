// extern "C" surface of libfastmat_b200.so (declared in include/fastmat_b200.h) and the plan objects behind it.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <new>

#include "common.h"
#include "fft_engine.h"

namespace fmb {

// ------------------------------------------------------------------------------------------- globals
static thread_local std::string t_error;
std::atomic<long long> g_launches(0);

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_error = buf;
}

const DeviceProps &device_props() {
    static DeviceProps props;
    static std::once_flag once;
    std::call_once(once, []() {
#ifdef FMB_EMULATE
        props.sm_count = 148; props.cc_major = 10; props.cc_minor = 0;
        props.l2_bytes = (size_t)126 << 20; props.smem_optin = 227 << 10; props.ok = true;
#else
        int dev = 0;
        cudaDeviceProp p;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&p, dev) == cudaSuccess) {
            props.sm_count = p.multiProcessorCount;
            props.cc_major = p.major; props.cc_minor = p.minor;
            props.l2_bytes = (size_t)p.l2CacheSize;
            props.smem_optin = p.sharedMemPerBlockOptin;
            props.ok = true;
        } else {
            cudaGetLastError();
        }
#endif
    });
    return props;
}

static int require_device() {
    if (!device_props().ok) {
        set_error("no CUDA device available: fastmat_b200 has no CPU fallback");
        return FMB_ERR_CUDA;
    }
    return FMB_OK;
}

// kernels implemented in the other translation units
int fwht_apply(int order, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs, int64_t ycs, int64_t M, int dtype, void *ws,
               int64_t ws_bytes, cudaStream_t st);
int diag_apply(const void *d_dev, int dt_d, int64_t n, int conj_d, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs,
               int64_t ycs, int64_t M, int dt_x, int dt_out, cudaStream_t st);
int gather_apply(const void *idx_dev, int64_t nsel, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs, int64_t ycs,
                 int64_t M, int dtype, cudaStream_t st);
int scatter_inverse_apply(const void *inv_dev, int64_t ntotal, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs,
                          int64_t ycs, int64_t M, int dtype, cudaStream_t st);
int scatter_apply(const void *idx_dev, int64_t nsel, int64_t ntotal, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs,
                  int64_t ycs, int64_t M, int dtype, cudaStream_t st);
int conj_apply(const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs, int64_t ycs, int64_t n, int64_t M, int dtype, cudaStream_t st);
int ista_step_apply(const void *x, const void *grad, void *step_out, void *x_out, int64_t count, double numL, double alpha, int dtype,
                    cudaStream_t st);
int cast_apply(const void *x, int64_t xrs, int64_t xcs, int dt_x, void *y, int64_t yrs, int64_t ycs, int dt_out, int64_t n, int64_t M,
               cudaStream_t st);
int gs_step_apply(int subtract, const void *q, int64_t q_bs, int64_t q_rs, int k, void *v, int64_t v_bs, int64_t n, int64_t batches,
                  void *coef, int64_t coef_bs, int dtype, cudaStream_t st);
int64_t abs_argmax_workspace_bytes(int64_t cols);
int abs_argmax_apply(const void *x, int64_t rows, int64_t cols, int64_t col_stride, int dtype, int64_t *out, void *ws, int64_t ws_bytes,
                     cudaStream_t st);

static size_t out_csize(int dt_out) { return dt_out == FMB_COMPLEX64 ? sizeof(float2) : sizeof(double2); }

// ------------------------------------------------------------------------------------------- engine-backed plans
struct EnginePlan : PlanBase {
    ConvEngine eng;
    // Fourier orders run as chirp-z over a power of two although they are directly transformable (make_fourier) keep the
    // direct engine too: complex128 has no specialised single kernel at a padded length of 4096 and is better served by it
    std::unique_ptr<ConvEngine> direct;
    bool use_direct(int dt_out) const { return direct && dt_out == FMB_COMPLEX128 && eng.L == 4096; }
    int64_t bluestein_ref = 0;        // the reference's _numL decision, for the record
    int info(fmb_plan_info *o) const override {
        memset(o, 0, sizeof(*o));
        o->kind = kind; o->num_rows = num_rows; o->num_cols = num_cols;
        o->inner_size = eng.L; o->bluestein = bluestein_ref; o->passes_fwd = eng.passes();
        {   // schedule of a large complex64 column batch (what bench.py reports)
            int cols = 0, ns = 1;
            eng.slab_plan(1 << 20, sizeof(float2), eng.shape.pow2, cols, ns);
            o->slab_cols = cols;
        }
        return FMB_OK;
    }
    int64_t workspace_bytes(int, int64_t M, int, int dt_out) const override {
        return use_direct(dt_out) ? direct->workspace_bytes(M, out_csize(dt_out)) : eng.workspace_bytes(M, out_csize(dt_out));
    }
    int apply(int direction, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs, int64_t ycs, int64_t M, int dt_in,
              int dt_out, void *ws, int64_t wsb, cudaStream_t st) const override {
        if (use_direct(dt_out)) return direct->run(direction, x, xrs, xcs, y, yrs, ycs, M, dt_in, dt_out, ws, wsb, st);
        return eng.run(direction, x, xrs, xcs, y, yrs, ycs, M, dt_in, dt_out, ws, wsb, st);
    }
};

// chirp-z (Bluestein) constants: fastmat/Fourier.pyx:127-156, with k^2 reduced mod 2N in integers first
static int setup_bluestein(ConvEngine &eng, int64_t N) {
    int64_t L = next_pow2(2 * N - 1);
    if (L < 2) L = 2;
    int rc = eng.init(L, N, N, true);
    if (rc) return rc;
    std::vector<cd> w((size_t)N), chirp((size_t)L, cd(0, 0)), chat;
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int64_t k = 0; k < N; ++k) {
        unsigned long long e = ((unsigned long long)k * (unsigned long long)k) % (unsigned long long)(2 * N);
        long double a = pi * (long double)e / (long double)N;
        w[(size_t)k] = cd((double)cosl(a), (double)-sinl(a));           // exp(-i pi k^2 / N)
    }
    for (int64_t k = 0; k < N; ++k) {
        chirp[(size_t)k] = std::conj(w[(size_t)k]);
        if (k > 0) chirp[(size_t)(L - k)] = std::conj(w[(size_t)k]);
    }
    if ((rc = device_fft_c128(chirp, chat))) return rc;
    for (auto &v : chat) v /= (double)L;                                // the 1/L of the inverse transform
    eng.pre = w; eng.post = w; eng.mid = chat;
    return FMB_OK;
}

static int make_fourier(fmb_plan **out, int64_t order, int optimize, int max_stage) {
    if (order < 1) { set_error("Fourier order cannot be smaller than 1."); return FMB_ERR_VALUE; }     // Fourier.pyx:100-101
    int rc = require_device();
    if (rc) return rc;
    std::unique_ptr<EnginePlan> p(new EnginePlan());
    p->kind = FMB_KIND_FOURIER; p->num_rows = order; p->num_cols = order;
    if (optimize) {                                                     // Fourier.pyx:109-122
        int64_t padded = find_optimal_fft_size(order * 2 - 1, max_stage);
        float rhs_f = 2 * fft_complexity(padded) + (float)(2 * padded);
        double rhs = (double)rhs_f + 2.0 * (double)order;
        p->bluestein_ref = ((double)fft_complexity(order) < rhs) ? 0 : padded;
    }
    // directly transformable lengths that are not powers of two would run the run-time-radix kernels (3 - 7 % of the
    // roofline); the chirp-z transform over the next power of two (specialised kernels: one launch up to a padded length
    // of 4096, the two-pass path above) is 1.5 - 6x faster from order 33 on (FMB_POW2_PAD=0: direct).  `bluestein_ref`
    // above keeps reporting the reference's own decision.
    static const long pow2_pad = getenv("FMB_POW2_PAD") ? atol(getenv("FMB_POW2_PAD")) : 1;
    FftShape shape;
    bool direct = plan_shape(order, shape);
    if (direct && pow2_pad && (order & (order - 1)) != 0 && order > 32 && next_pow2(2 * order - 1) <= ((int64_t)1 << 24)) direct = false;
    if (direct) rc = p->eng.init(order, order, order, false);
    else {
        rc = setup_bluestein(p->eng, order);
        if (rc == FMB_OK && p->eng.L == 4096 && plan_shape(order, shape)) {
            p->direct.reset(new ConvEngine());
            rc = p->direct->init(order, order, order, false);
        }
    }
    if (rc) return rc;
    *out = reinterpret_cast<fmb_plan *>(static_cast<PlanBase *>(p.release()));
    return FMB_OK;
}

// circular-convolution plan shared by Circulant and Toeplitz: first column `gen` of the L x L circulant
static int make_conv(fmb_plan **out, int kind, const std::vector<cd> &gen, int64_t L, int64_t n_in, int64_t n_out) {
    std::unique_ptr<EnginePlan> p(new EnginePlan());
    p->kind = kind; p->num_rows = n_out; p->num_cols = n_in;
    int rc = p->eng.init(L, n_in, n_out, true);
    if (rc) return rc;
    std::vector<cd> spec;
    if ((rc = device_fft_c128(gen, spec))) return rc;
    for (auto &v : spec) v /= (double)L;                                // Circulant.pyx:131, Toeplitz.pyx:279
    p->eng.mid = spec;
    *out = reinterpret_cast<fmb_plan *>(static_cast<PlanBase *>(p.release()));
    return FMB_OK;
}

// The reference pads to the length its CPU cost model likes best (2^a 3^b 5^c ..., cmath.pyx:88-153); any length that
// embeds the operator gives the same result.  On this device the power-of-two lengths run compile-time specialised kernels
// at 25 - 90 % of the HBM roofline and every other length the run-time-radix kernels at 3 - 7 %, so a non-power-of-two
// choice is replaced by the next power of two that still embeds the operator (`need`), unless that is out of the
// specialised range (FMB_POW2_PAD=0: keep the reference's choice).  The planner functions themselves stay bit-identical.
static int64_t prefer_pow2_length(int64_t chosen, int64_t /*valid_exact*/, int64_t need) {
    static const long on = getenv("FMB_POW2_PAD") ? atol(getenv("FMB_POW2_PAD")) : 1;
    if (!on || chosen < 2 || (chosen & (chosen - 1)) == 0) return chosen;
    const int64_t p2 = next_pow2(need);
    if (p2 < 64 || p2 > ((int64_t)1 << 24)) return chosen;
    return p2;
}

static int make_circulant(fmb_plan **out, const void *c_host, int64_t n, int optimize, int max_stage) {
    if (n < 1 || !c_host) { set_error("Column-definition tensor must be at least 1D."); return FMB_ERR_VALUE; }
    int rc = require_device();
    if (rc) return rc;
    const cd *c = (const cd *)c_host;
    int64_t L = n;                                                      // Circulant.pyx:104-124
    if (optimize) {
        int64_t padded = find_optimal_fft_size(2 * n - 1, max_stage);
        if (fft_complexity(n) > fft_complexity(padded)) L = padded;
    }
    FftShape shape;
    if (!plan_shape(L, shape) || (L != n && L < 2 * n - 1)) L = std::max<int64_t>(2, next_pow2(2 * n - 1));
    L = prefer_pow2_length(L, L == n ? n : 2 * n - 1, 2 * n - 1);
    std::vector<cd> gen((size_t)L, cd(0, 0));
    for (int64_t i = 0; i < n; ++i) gen[(size_t)i] = c[i];
    if (L != n)
        for (int64_t i = 1; i < n; ++i) gen[(size_t)(L - n + i)] = c[i];   // [c, 0..., c[1:]]
    return make_conv(out, FMB_KIND_CIRCULANT, gen, L, n, n);
}

static int make_toeplitz(fmb_plan **out, const void *vc_host, int64_t n, const void *vr_host, int64_t m1, int optimize, int max_stage) {
    if (n < 1 || m1 < 0 || !vc_host || (m1 > 0 && !vr_host)) {
        set_error("Column- and row-definition vectors must be 1D.");
        return FMB_ERR_VALUE;
    }
    int rc = require_device();
    if (rc) return rc;
    const cd *vc = (const cd *)vc_host, *vr = (const cd *)vr_host;
    const int64_t m = m1 + 1, d = n + m - 1;
    int64_t L = d;                                                      // Toeplitz.pyx:225-233
    if (optimize) {
        int64_t opt = find_optimal_fft_size(d, max_stage);
        if (fft_complexity(opt) < fft_complexity(d)) L = opt;
    }
    FftShape shape;
    if (!plan_shape(L, shape) || L < d) L = std::max<int64_t>(2, next_pow2(d));
    L = prefer_pow2_length(L, d, d);
    std::vector<cd> gen((size_t)L, cd(0, 0));                           // [vecC, zeros, vecR]  (_preProcSlice :357-366)
    for (int64_t i = 0; i < n; ++i) gen[(size_t)i] = vc[i];
    for (int64_t i = 0; i < m1; ++i) gen[(size_t)(L - m1 + i)] = vr[i];
    return make_conv(out, FMB_KIND_TOEPLITZ, gen, L, m, n);
}

static int make_kron_fourier(fmb_plan **out, const int64_t *dims, int ndims) {
    if (ndims < 2 || !dims) { set_error("Kronecker: Product must have at least two terms"); return FMB_ERR_VALUE; }   // Kron.pyx:100-101
    if (ndims != 2) { set_error("Kron(Fourier...) with %d factors is composed by the class layer", ndims); return FMB_ERR_NOTIMPL; }
    if (dims[0] < 1 || dims[1] < 1) { set_error("Fourier order cannot be smaller than 1."); return FMB_ERR_VALUE; }
    int rc = require_device();
    if (rc) return rc;
    std::unique_ptr<EnginePlan> p(new EnginePlan());
    p->kind = FMB_KIND_KRON_FOURIER; p->num_rows = p->num_cols = dims[0] * dims[1];
    if ((rc = p->eng.init_kron(dims[0], dims[1]))) return rc;
    *out = reinterpret_cast<fmb_plan *>(static_cast<PlanBase *>(p.release()));
    return FMB_OK;
}

// ------------------------------------------------------------------------------------------- Hadamard / Diag / Partial
struct HadamardPlan : PlanBase {
    int order = 0;
    // per-slab counters of the persistent order-20 kernel (experiment, FMB_FWHT_PERSIST): at most one slab per column
    int64_t workspace_bytes(int, int64_t M, int, int) const override { return order == 20 ? (M + 64) * 4 : 0; }
    int info(fmb_plan_info *o) const override {
        memset(o, 0, sizeof(*o));
        o->kind = kind; o->num_rows = num_rows; o->num_cols = num_cols; o->inner_size = num_rows; o->passes_fwd = order > 12 ? 2 : 1;
        return FMB_OK;
    }
    int apply(int, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs, int64_t ycs, int64_t M, int dt_in, int dt_out,
              void *ws, int64_t ws_bytes, cudaStream_t st) const override {
        if (dt_in != dt_out) { set_error("Hadamard: output dtype must equal input dtype (promote(in, int8) = in)"); return FMB_ERR_TYPE; }
        return fwht_apply(order, x, xrs, xcs, y, yrs, ycs, M, dt_in, ws, ws_bytes, st);   // symmetric: backward == forward (:232-239)
    }
};

struct DiagPlan : PlanBase {
    DevArray d;
    int dt = 0;
    int info(fmb_plan_info *o) const override {
        memset(o, 0, sizeof(*o));
        o->kind = kind; o->num_rows = num_rows; o->num_cols = num_cols; o->inner_size = num_rows; o->passes_fwd = 1;
        return FMB_OK;
    }
    int apply(int direction, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs, int64_t ycs, int64_t M, int dt_in,
              int dt_out, void *, int64_t, cudaStream_t st) const override {
        return diag_apply(d.p, dt, num_rows, direction == FMB_BACKWARD, x, xrs, xcs, y, yrs, ycs, M, dt_in, dt_out, st);
    }
};

struct PartialPlan : PlanBase {
    DevArray idx;
    DevArray sidx;           // scatter targets without an inverse table: idx with all but the last occurrence of a value set to -1
    DevArray inv;            // inverse index (ntotal entries, -1 = row not selected), absent for very sparse selections
    int64_t nsel = 0, ntotal = 0;
    int info(fmb_plan_info *o) const override {
        memset(o, 0, sizeof(*o));
        o->kind = kind; o->num_rows = num_rows; o->num_cols = num_cols; o->inner_size = ntotal; o->passes_fwd = 1;
        return FMB_OK;
    }
    int apply(int direction, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs, int64_t ycs, int64_t M, int dt_in,
              int dt_out, void *, int64_t, cudaStream_t st) const override {
        if (dt_in != dt_out) { set_error("Partial: gather/scatter does not convert dtypes"); return FMB_ERR_TYPE; }
        if (direction == FMB_FORWARD) return gather_apply(idx.p, nsel, x, xrs, xcs, y, yrs, ycs, M, dt_in, st);
        if (inv.p) return scatter_inverse_apply(inv.p, ntotal, x, xrs, xcs, y, yrs, ycs, M, dt_in, st);
        return scatter_apply(sidx.p ? sidx.p : idx.p, nsel, ntotal, x, xrs, xcs, y, yrs, ycs, M, dt_in, st);
    }
};

}  // namespace fmb

using namespace fmb;

#define FMB_GUARD_BEGIN try {
#define FMB_GUARD_END                                              \
    }                                                              \
    catch (const std::bad_alloc &) {                               \
        set_error("out of host memory");                           \
        return FMB_ERR_CUDA;                                       \
    }                                                              \
    catch (const std::exception &e) {                              \
        set_error("internal error: %s", e.what());                 \
        return FMB_ERR_CUDA;                                       \
    }

extern "C" {

const char *fmb_last_error(void) { return t_error.c_str(); }
int fmb_version(void) { return 100; }

int fmb_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *l2_bytes) {
    int rc = require_device();
    if (rc) return rc;
    const DeviceProps &p = device_props();
    if (sm_count) *sm_count = p.sm_count;
    if (cc_major) *cc_major = p.cc_major;
    if (cc_minor) *cc_minor = p.cc_minor;
    if (l2_bytes) *l2_bytes = p.l2_bytes;
    return FMB_OK;
}

int64_t fmb_find_optimal_fft_size(int64_t order, int max_stage) { return find_optimal_fft_size(order, max_stage); }
float fmb_fft_complexity(int64_t n) { return fft_complexity(n); }

int fmb_lfsr_order(uint32_t polynomial) { return lfsr_order(polynomial); }

int64_t fmb_lfsr_period(uint32_t polynomial, uint32_t start) {
    const int64_t p = lfsr_period(polynomial, start);
    if (p == -1) set_error("Only polynomials of order 1 to 31 are supported.");          // LFSRCirculant.pyx:201-202
    else if (p == -2) set_error("Initial state must be non-zero.");                       // :204-205
    else if (p == -3) set_error("Register configuration produces invalid sequence.");     // :218-220
    return p < 0 ? (int64_t)FMB_ERR_VALUE : p;
}

int fmb_lfsr_sequences(uint32_t polynomial, uint32_t start, int64_t n, uint32_t *gen_states, uint32_t *tap_states,
                       int8_t *vec_c) {
    const int order = lfsr_order(polynomial);
    if (order > 31 || order < 1) { set_error("Only polynomials of order 1 to 31 are supported."); return FMB_ERR_VALUE; }
    if (n < 0) { set_error("LFSR: negative sequence length"); return FMB_ERR_VALUE; }
    lfsr_sequences(polynomial, start, n, gen_states, tap_states, vec_c);
    return FMB_OK;
}

int fmb_fourier_plan_create(fmb_plan **out, int64_t order, int optimize, int max_stage) {
    FMB_GUARD_BEGIN
    if (!out) { set_error("null output pointer"); return FMB_ERR_VALUE; }
    return make_fourier(out, order, optimize, max_stage);
    FMB_GUARD_END
}

int fmb_circulant_plan_create(fmb_plan **out, const void *c_host, int64_t n, int optimize, int max_stage) {
    FMB_GUARD_BEGIN
    if (!out) { set_error("null output pointer"); return FMB_ERR_VALUE; }
    return make_circulant(out, c_host, n, optimize, max_stage);
    FMB_GUARD_END
}

int fmb_toeplitz_plan_create(fmb_plan **out, const void *vec_c_host, int64_t n, const void *vec_r_host, int64_t m_minus_1,
                             int optimize, int max_stage) {
    FMB_GUARD_BEGIN
    if (!out) { set_error("null output pointer"); return FMB_ERR_VALUE; }
    return make_toeplitz(out, vec_c_host, n, vec_r_host, m_minus_1, optimize, max_stage);
    FMB_GUARD_END
}

int fmb_hadamard_plan_create(fmb_plan **out, int order) {
    FMB_GUARD_BEGIN
    if (!out) { set_error("null output pointer"); return FMB_ERR_VALUE; }
    if (order < 1) { set_error("Hadamard: Order must be larger than 0."); return FMB_ERR_VALUE; }          // Hadamard.pyx:116-117
    if (order > 62) { set_error("Hadamard: Order exceeds maximum for this platform: %d", 62); return FMB_ERR_VALUE; }  // :119-123
    int rc = require_device();
    if (rc) return rc;
    HadamardPlan *p = new HadamardPlan();
    p->kind = FMB_KIND_HADAMARD; p->order = order; p->num_rows = p->num_cols = (int64_t)1 << order;
    *out = reinterpret_cast<fmb_plan *>(static_cast<PlanBase *>(p));
    return FMB_OK;
    FMB_GUARD_END
}

int fmb_diag_plan_create(fmb_plan **out, const void *d_host, int dtype, int64_t n) {
    FMB_GUARD_BEGIN
    if (!out) { set_error("null output pointer"); return FMB_ERR_VALUE; }
    if (n < 1 || !d_host) { set_error("Diag: Definition vector must have exactly one dimension."); return FMB_ERR_VALUE; }
    if (dtype_size(dtype) == 0) { set_error("Diag: unsupported dtype %d", dtype); return FMB_ERR_TYPE; }
    int rc = require_device();
    if (rc) return rc;
    std::unique_ptr<DiagPlan> p(new DiagPlan());
    p->kind = FMB_KIND_DIAG; p->num_rows = p->num_cols = n; p->dt = dtype;
    if ((rc = p->d.upload(d_host, (size_t)n * dtype_size(dtype)))) return rc;
    *out = reinterpret_cast<fmb_plan *>(static_cast<PlanBase *>(p.release()));
    return FMB_OK;
    FMB_GUARD_END
}

int fmb_partial_plan_create(fmb_plan **out, const int64_t *idx_host, int64_t num_sel, int64_t num_total) {
    FMB_GUARD_BEGIN
    if (!out) { set_error("null output pointer"); return FMB_ERR_VALUE; }
    if (num_sel < 0 || num_total < 1 || (num_sel > 0 && !idx_host)) { set_error("Partial: bad selection"); return FMB_ERR_VALUE; }
    for (int64_t i = 0; i < num_sel; ++i)
        if (idx_host[i] < 0 || idx_host[i] >= num_total) {
            set_error("Partial: An index exceeds matrix dimensions.");                                       // Partial.pyx:158-162
            return FMB_ERR_VALUE;
        }
    int rc = require_device();
    if (rc) return rc;
    std::unique_ptr<PartialPlan> p(new PartialPlan());
    p->kind = FMB_KIND_PARTIAL; p->num_rows = num_sel; p->num_cols = num_total; p->nsel = num_sel; p->ntotal = num_total;
    if ((rc = p->idx.upload(idx_host, (size_t)num_sel * sizeof(int64_t)))) return rc;
    // The backward direction (y = 0; y[idx[r]] = x[r]) runs as a gather through the inverse index: no zero-fill pass,
    // coalesced stores, and a repeated index resolves like numpy's assignment (the last occurrence wins) instead of
    // racing.  Skipped when the table would dwarf the selection (a few rows out of a huge matrix).
    static const long inv_on = getenv("FMB_PARTIAL_INVERSE") ? atol(getenv("FMB_PARTIAL_INVERSE")) : 1;
    if (inv_on && num_sel > 0 && (num_total <= (int64_t)(32 << 20) || num_sel * 8 >= num_total)) {
        std::vector<int64_t> inv((size_t)num_total, (int64_t)-1);
        for (int64_t i = 0; i < num_sel; ++i) inv[(size_t)idx_host[i]] = i;
        if ((rc = p->inv.upload(inv.data(), (size_t)num_total * sizeof(int64_t)))) return rc;
    } else if (num_sel > 0) {
        // zero + scatter path: keep only the LAST occurrence of a repeated index (numpy semantics, deterministic)
        std::vector<int64_t> order((size_t)num_sel), sidx(idx_host, idx_host + num_sel);
        for (int64_t i = 0; i < num_sel; ++i) order[(size_t)i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return idx_host[a] < idx_host[b]; });
        bool repeats = false;
        for (int64_t k = 0; k + 1 < num_sel; ++k)
            if (idx_host[order[(size_t)k]] == idx_host[order[(size_t)k + 1]]) { sidx[(size_t)order[(size_t)k]] = -1; repeats = true; }
        if (repeats && (rc = p->sidx.upload(sidx.data(), (size_t)num_sel * sizeof(int64_t)))) return rc;
    }
    *out = reinterpret_cast<fmb_plan *>(static_cast<PlanBase *>(p.release()));
    return FMB_OK;
    FMB_GUARD_END
}

int fmb_kron_fourier_plan_create(fmb_plan **out, const int64_t *dims, int ndims) {
    FMB_GUARD_BEGIN
    if (!out) { set_error("null output pointer"); return FMB_ERR_VALUE; }
    return make_kron_fourier(out, dims, ndims);
    FMB_GUARD_END
}

int fmb_plan_info_get(const fmb_plan *plan, fmb_plan_info *info) {
    if (!plan || !info) { set_error("null pointer"); return FMB_ERR_VALUE; }
    return reinterpret_cast<const PlanBase *>(plan)->info(info);
}

int64_t fmb_plan_workspace_bytes(const fmb_plan *plan, int direction, int64_t M, int dtype_in, int dtype_out) {
    if (!plan) return 0;
    return reinterpret_cast<const PlanBase *>(plan)->workspace_bytes(direction, M, dtype_in, dtype_out);
}

int fmb_plan_apply(const fmb_plan *plan, int direction, const void *x, int64_t x_row_stride, int64_t x_col_stride, void *y,
                   int64_t y_row_stride, int64_t y_col_stride, int64_t M, int dtype_in, int dtype_out, void *workspace,
                   int64_t workspace_bytes, void *cuda_stream) {
    FMB_GUARD_BEGIN
    if (!plan) { set_error("null plan"); return FMB_ERR_VALUE; }
    if (direction != FMB_FORWARD && direction != FMB_BACKWARD) { set_error("bad direction %d", direction); return FMB_ERR_VALUE; }
    if (M < 0) { set_error("negative column count"); return FMB_ERR_VALUE; }
    if (M == 0) return FMB_OK;
    if (!x || !y) { set_error("null data pointer"); return FMB_ERR_VALUE; }
    if (x == y) { set_error("x and y must not alias"); return FMB_ERR_VALUE; }
    return reinterpret_cast<const PlanBase *>(plan)->apply(direction, x, x_row_stride, x_col_stride, y, y_row_stride, y_col_stride, M,
                                                          dtype_in, dtype_out, workspace, workspace_bytes, (cudaStream_t)cuda_stream);
    FMB_GUARD_END
}

int fmb_plan_destroy(fmb_plan *plan) {
    if (plan) delete reinterpret_cast<PlanBase *>(plan);
    return FMB_OK;
}

int fmb_conjugate(const void *x, int64_t x_row_stride, int64_t x_col_stride, void *y, int64_t y_row_stride, int64_t y_col_stride,
                  int64_t n, int64_t M, int dtype, void *cuda_stream) {
    FMB_GUARD_BEGIN
    return conj_apply(x, x_row_stride, x_col_stride, y, y_row_stride, y_col_stride, n, M, dtype, (cudaStream_t)cuda_stream);
    FMB_GUARD_END
}

int fmb_cast(const void *x, int64_t x_row_stride, int64_t x_col_stride, int dtype_in, void *y, int64_t y_row_stride,
             int64_t y_col_stride, int dtype_out, int64_t n, int64_t M, void *cuda_stream) {
    FMB_GUARD_BEGIN
    return cast_apply(x, x_row_stride, x_col_stride, dtype_in, y, y_row_stride, y_col_stride, dtype_out, n, M, (cudaStream_t)cuda_stream);
    FMB_GUARD_END
}

int fmb_ista_step(const void *x, const void *grad, void *step_out, void *x_out, int64_t count, double num_l, double alpha, int dtype,
                  void *cuda_stream) {
    FMB_GUARD_BEGIN
    return ista_step_apply(x, grad, step_out, x_out, count, num_l, alpha, dtype, (cudaStream_t)cuda_stream);
    FMB_GUARD_END
}

int64_t fmb_abs_argmax_workspace_bytes(int64_t cols) { return abs_argmax_workspace_bytes(cols); }

int fmb_abs_argmax(const void *x, int64_t rows, int64_t cols, int64_t col_stride, int dtype, int64_t *out_index, void *workspace,
                   int64_t workspace_bytes, void *cuda_stream) {
    FMB_GUARD_BEGIN
    return abs_argmax_apply(x, rows, cols, col_stride, dtype, out_index, workspace, workspace_bytes, (cudaStream_t)cuda_stream);
    FMB_GUARD_END
}

int fmb_gs_project(const void *q, int64_t q_batch_stride, int64_t q_row_stride, int k, const void *v, int64_t v_batch_stride, int64_t n,
                   int64_t batches, void *coef, int64_t coef_batch_stride, int dtype, void *cuda_stream) {
    FMB_GUARD_BEGIN
    return gs_step_apply(0, q, q_batch_stride, q_row_stride, k, const_cast<void *>(v), v_batch_stride, n, batches, coef, coef_batch_stride,
                         dtype, (cudaStream_t)cuda_stream);
    FMB_GUARD_END
}

int fmb_gs_subtract(const void *q, int64_t q_batch_stride, int64_t q_row_stride, int k, void *v, int64_t v_batch_stride, int64_t n,
                    int64_t batches, const void *coef, int64_t coef_batch_stride, int dtype, void *cuda_stream) {
    FMB_GUARD_BEGIN
    return gs_step_apply(1, q, q_batch_stride, q_row_stride, k, v, v_batch_stride, n, batches, const_cast<void *>(coef), coef_batch_stride,
                         dtype, (cudaStream_t)cuda_stream);
    FMB_GUARD_END
}

int64_t fmb_launch_count(void) { return g_launches.load(); }

}  // extern "C"

#ifdef V32_TIMING
// experiment builds only (-DV32_TIMING): print the phase timings collected by the instrumented kernels
namespace fmb { std::vector<void (*)()> &debug_dumpers() { static std::vector<void (*)()> v; return v; } }
extern "C" void fmb_debug_dump(void) { for (auto f : fmb::debug_dumpers()) f(); }
#endif
