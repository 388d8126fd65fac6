// Host side of the FFT / convolution engine: shape planning, constant tables, pass construction, launches.
#include "fft_engine.h"

#include <algorithm>
#include <cmath>

#include "fft_fast.cuh"
#include "fft_v32.cuh"
#include "fft_v32p.cuh"
#include "fft_pass.cuh"

namespace fmb {

// ------------------------------------------------------------------------------------------- kernels
// one translation unit per (precision, pow2) instantiation: fft_k_*.cu
int launch_fft_f32_pow2(const PassParams<float2> &p, unsigned tiles, int nt, size_t smem, cudaStream_t st);
int launch_fft_f32_gen(const PassParams<float2> &p, unsigned tiles, int nt, size_t smem, cudaStream_t st);
int launch_fft_f64_pow2(const PassParams<double2> &p, unsigned tiles, int nt, size_t smem, cudaStream_t st);
int launch_fft_f64_gen(const PassParams<double2> &p, unsigned tiles, int nt, size_t smem, cudaStream_t st);
static int launch_fft_kernel(const PassParams<float2> &p, bool pow2, unsigned tiles, int nt, size_t smem, cudaStream_t st) {
    return pow2 ? launch_fft_f32_pow2(p, tiles, nt, smem, st) : launch_fft_f32_gen(p, tiles, nt, smem, st);
}
static int launch_fft_kernel(const PassParams<double2> &p, bool pow2, unsigned tiles, int nt, size_t smem, cudaStream_t st) {
    return pow2 ? launch_fft_f64_pow2(p, tiles, nt, smem, st) : launch_fft_f64_gen(p, tiles, nt, smem, st);
}

static long env_long(const char *name, long dflt) {
    const char *v = getenv(name);
    return v ? atol(v) : dflt;
}

// ------------------------------------------------------------------------------------------- shape planning
int64_t next_pow2(int64_t v) {
    int64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

static bool factor_small(int64_t n, std::vector<int> &primes) {
    static const int ps[] = {2, 3, 5, 7, 11, 13};
    primes.clear();
    for (int p : ps)
        while (n % p == 0) { primes.push_back(p); n /= p; }
    return n == 1;
}

// radices for one shared-memory pass of length R (R's primes all <= 13)
static void make_radices(int64_t R, PassGeom &g) {
    std::vector<int> primes;
    factor_small(R, primes);
    int twos = 0;
    g.radix.clear();
    g.R = (int)R;
    for (int p : primes) {
        if (p == 2) ++twos;
        else g.radix.push_back(p);
    }
    std::sort(g.radix.begin(), g.radix.end(), [](int a, int b) { return a > b; });
    // powers of two: as many radix-16 stages as possible, remainder 8 / 4 / 2 placed first
    std::vector<int> two_r;
    while (twos >= 4) { two_r.push_back(16); twos -= 4; }
    if (twos == 3) two_r.insert(two_r.begin(), 8);
    else if (twos == 2) two_r.insert(two_r.begin(), 4);
    else if (twos == 1) two_r.insert(two_r.begin(), 2);
    g.radix.insert(g.radix.end(), two_r.begin(), two_r.end());
    if (g.radix.empty()) g.radix.push_back(1);        // R == 1 handled by caller (never planned)
    g.min_pnb = 16;
    for (int p : g.radix) g.min_pnb = std::min(g.min_pnb, p * (FMB_EMAX / p));
}

static int64_t pass_limit(const PassGeom &g) { return (int64_t)FMB_MAX_NT * g.min_pnb; }   // R*T <= limit, T >= 1

bool plan_shape(int64_t L, FftShape &s, bool prefer_two) {
    s = FftShape();
    s.L = L;
    if (L < 2) return false;
    std::vector<int> primes;
    if (!factor_small(L, primes)) return false;
    s.pow2 = (L & (L - 1)) == 0;
    if (primes.size() > (size_t)2 * FMB_MAX_STAGES) return false;
    PassGeom g;
    make_radices(L, g);
    if (L <= pass_limit(g) && (int)g.radix.size() <= FMB_MAX_STAGES && L <= 8192 &&
        !(prefer_two && (L & (L - 1)) == 0 && L > 4096)) {
        s.npass = 1;
        s.g[0] = g;
        return true;
    }
    // two passes: L = R1 * R2, balanced
    int64_t R1 = 1, R2 = 1;
    if (s.pow2) {
        int lg = 0;
        while (((int64_t)1 << lg) < L) ++lg;
        R1 = (int64_t)1 << (lg / 2);
        static const long split1 = env_long("FMB_SPLIT_LOG1", 0);      // experiments: log2 of the first pass length
        if (split1 >= 8 && split1 <= 12 && lg - split1 >= 8 && lg - split1 <= 12) R1 = (int64_t)1 << split1;
        R2 = L / R1;
    } else {
        std::sort(primes.begin(), primes.end(), [](int a, int b) { return a > b; });
        for (int p : primes) {
            if (R1 <= R2) R1 *= p;
            else R2 *= p;
        }
        if (R1 > R2) std::swap(R1, R2);
    }
    if (R1 < 2 || R2 < 2) return false;
    make_radices(R1, s.g[0]);
    make_radices(R2, s.g[1]);
    if (R1 > pass_limit(s.g[0]) || R2 > pass_limit(s.g[1]) || R1 > 8192 || R2 > 8192) return false;
    if ((int)s.g[0].radix.size() > FMB_MAX_STAGES || (int)s.g[1].radix.size() > FMB_MAX_STAGES) return false;
    s.npass = 2;
    return true;
}

// ------------------------------------------------------------------------------------------- engine setup
int ConvEngine::init(int64_t L_, int64_t n_in_, int64_t n_out_, bool two_ffts_) {
    L = L_; n_in = n_in_; n_out = n_out_; two_ffts = two_ffts_;
    // 8192 = 64 x 128: two specialised passes over an L2-resident intermediate beat the one generic kernel (5x)
#ifdef FMB_EMULATE
    const bool two = false;
#else
    static const long single8k = env_long("FMB_SINGLE_8192", 0);
    const bool two = !single8k;
#endif
    if (!plan_shape(L, shape, two)) {
        set_error("FFT length %lld is not directly transformable", (long long)L);
        return FMB_ERR_VALUE;
    }
    return FMB_OK;
}

int ConvEngine::init_kron(int64_t a, int64_t b) {
    kron_a = a; kron_b = b;
    L = a * b; n_in = L; n_out = L; two_ffts = false;
    FftShape sa, sb;
    if (!plan_shape(a, sa) || !plan_shape(b, sb) || sa.npass != 1 || sb.npass != 1) {
        set_error("Kron(Fourier(%lld), Fourier(%lld)): factor sizes must be single-pass transformable", (long long)a, (long long)b);
        return FMB_ERR_NOTIMPL;
    }
    shape = FftShape();
    shape.L = L;
    shape.npass = 2;
    shape.g[0] = sa.g[0];
    shape.g[1] = sb.g[0];
    shape.pow2 = sa.pow2 && sb.pow2;
    return FMB_OK;
}

template <typename C> static int upload_cvec(DevArray &d, const std::vector<cd> &h) {
    typedef typename real_of<C>::type S;
    std::vector<C> tmp(h.size());
    for (size_t i = 0; i < h.size(); ++i) { tmp[i].x = (S)h[i].real(); tmp[i].y = (S)h[i].imag(); }
    return d.upload(tmp.data(), tmp.size() * sizeof(C));
}

static void unit_roots(int64_t n, int64_t count, int64_t stride, std::vector<cd> &w) {
    // w[j] = exp(-2 pi i (j*stride) / n), exact octant reduction for accuracy
    w.resize((size_t)count);
    const long double tp = 6.283185307179586476925286766559005768L;
    for (int64_t j = 0; j < count; ++j) {
        int64_t e = (j * stride) % n;
        long double a = tp * (long double)e / (long double)n;
        w[(size_t)j] = cd((double)cosl(a), (double)-sinl(a));
    }
}

static cd unit_root(int64_t e, int64_t n) {      // exp(-2 pi i e / n)
    const long double tp = 6.283185307179586476925286766559005768L;
    e %= n;
    long double a = tp * (long double)e / (long double)n;
    return cd((double)cosl(a), (double)-sinl(a));
}

template <typename C> int ConvEngine::ensure_dev(Dev &d) const {
    std::lock_guard<std::mutex> lock(mu);
    if (d.ready) return FMB_OK;
    std::vector<cd> w;
    for (int g = 0; g < shape.npass; ++g) {
        unit_roots(shape.g[g].R, shape.g[g].R, 1, w);
        int rc = upload_cvec<C>(d.wR[g], w);
        if (rc) return rc;
    }
    if (shape.pow2) {
        // fast path (fft_fast.cuh): per-stage twiddle tables of a pass length R = 16 * 16 * P2, pairs (r = 2p, 2p+1),
        // butterfly index fastest:  stage 2 (Ns = 16, radix 16) then stage 3 (Ns = 256, radix P2)
        for (int g = 0; g < shape.npass; ++g) {
            const int64_t Rg = shape.g[g].R;
            if (Rg < 64 || Rg > 4096) continue;
            auto root = unit_root;
            w.clear();
            const int P1 = Rg >= 256 ? 16 : (int)(Rg / 16);         // R = 128: radix 16 then radix 8 (FastPlan<7>)
            for (int p2 = 0; p2 < P1 / 2; ++p2)
                for (int kk = 0; kk < 16; ++kk) { w.push_back(root(kk * 2 * p2, 16 * P1)); w.push_back(root(kk * (2 * p2 + 1), 16 * P1)); }
            const int P2 = (int)(Rg / 256);
            for (int p2 = 0; p2 < P2 / 2; ++p2)
                for (int kk = 0; kk < 256; ++kk) { w.push_back(root((int64_t)kk * 2 * p2, Rg)); w.push_back(root((int64_t)kk * (2 * p2 + 1), Rg)); }
            int rc = upload_cvec<C>(d.twF[g], w);
            if (rc) return rc;
            if (Rg == 1024 && (shape.npass == 2 || two_ffts)) {
                // 32-values-per-thread passes (fft_v32.cuh): second-stage twiddles {W_1024^{kk 2p}, W_1024^{kk (2p+1)}} at
                // [p * 32 + kk], stored twice (see v32_pass_kernel), and the four-step step factor W_L^{32 i}
                w.clear();
                for (int copy = 0; copy < 2; ++copy)
                    for (int p2 = 0; p2 < 16; ++p2)
                        for (int kk = 0; kk < 32; ++kk) { w.push_back(root(kk * 2 * p2, 1024)); w.push_back(root(kk * (2 * p2 + 1), 1024)); }
                if ((rc = upload_cvec<C>(d.twV[g], w))) return rc;
                if (kron_a == 0 && shape.npass == 2) {
                    unit_roots(L, L / Rg, 32, w);
                    if ((rc = upload_cvec<C>(d.twS32[g], w))) return rc;
                }
            }
        }
    }
    if (shape.npass == 2 && kron_a == 0) {
        int lg = 0;
        while (((int64_t)1 << (2 * lg)) < L) ++lg;       // B = 2^lg >= sqrt(L)
        d.tw_shift = lg;
        int64_t B = (int64_t)1 << lg;
        unit_roots(L, B, 1, w);
        int rc = upload_cvec<C>(d.twL, w);
        if (rc) return rc;
        unit_roots(L, (L + B - 1) / B, B, w);
        rc = upload_cvec<C>(d.twH, w);
        if (rc) return rc;
        // fast path: twS[g][i] = W_L^{(R_g / 16) i}, i < L / R_g (step of the four-step twiddle along a thread's outputs)
        for (int g = 0; g < 2; ++g) {
            const int64_t Rg = shape.g[g].R;
            if (Rg % 16) continue;
            unit_roots(L, L / Rg, Rg / 16, w);
            if ((rc = upload_cvec<C>(d.twS[g], w))) return rc;
        }
    }
    int rc;
    if (!pre.empty() && (rc = upload_cvec<C>(d.pre, pre))) return rc;
    if (!post.empty() && (rc = upload_cvec<C>(d.post, post))) return rc;
    if (!mid.empty()) {
        if (shape.npass == 2) {
            // two-pass convolutions read the spectrum in the middle pass, whose lines are k1 and whose transform index is
            // k2 (X index k1 + R1 k2): store it transposed, midT[k1*R2 + k2], so that those reads are contiguous
            const int64_t R1 = shape.g[0].R, R2 = shape.g[1].R;
            std::vector<cd> mt(mid.size());
            for (int64_t k2 = 0; k2 < R2; ++k2)
                for (int64_t k1 = 0; k1 < R1; ++k1) mt[(size_t)(k1 * R2 + k2)] = mid[(size_t)(k2 * R1 + k1)];
            if ((rc = upload_cvec<C>(d.mid, mt))) return rc;
        } else if ((rc = upload_cvec<C>(d.mid, mid))) return rc;
    }
    d.ready = true;
    return FMB_OK;
}

// Columns pushed through all passes together, and how many such slabs are in flight on different streams.
//  * default for the fast path ("pipelined slabs"): slabs of a few columns whose intermediate stays in the 126 MB L2
//    between the passes, issued round-robin on `ns` internal streams so that the tail of one launch overlaps the head of
//    the next; HBM then sees only x and y (FMB_PIPE_STREAMS / FMB_PIPE_MB tune it, FMB_PIPE_STREAMS=1 switches it off);
//  * otherwise one launch per pass over a 512 MiB slab (FMB_SLAB_MB), intermediate through HBM.
void ConvEngine::slab_plan(int64_t M, size_t csize, bool fast, int &cols, int &ns) const {
    ns = 1;
    if (shape.npass == 1) { cols = (int)std::min<int64_t>(M, 1 << 30); return; }
    static const long slab_mb = env_long("FMB_SLAB_MB", 0), pipe_ns = env_long("FMB_PIPE_STREAMS", 3),
                      pipe_mb = env_long("FMB_PIPE_MB", 16);
    const size_t col_bytes = (size_t)L * csize;
    if (fast && pipe_ns > 1 && slab_mb == 0) {
        int64_t c = std::max<int64_t>(1, (int64_t)(((size_t)pipe_mb << 20) / col_bytes));
        if (M >= 2 * c * pipe_ns) { cols = (int)c; ns = (int)std::min<long>(pipe_ns, FMB_MAX_PIPE); return; }
    }
    size_t budget = (size_t)512 << 20;
    if (slab_mb > 0) budget = (size_t)slab_mb << 20;
    int64_t s = (int64_t)(budget / col_bytes);
    if (s < 1) s = 1;
    cols = (int)std::min<int64_t>(s, M);
}

int ConvEngine::slab_cols(int64_t M, size_t csize) const {
    int cols, ns;
    slab_plan(M, csize, false, cols, ns);
    return cols;
}

int transpose_apply(const void *in, int64_t lda, void *out, int64_t ldb, int64_t ni, int64_t nj, size_t esize, cudaStream_t st);
// columns per transposed chunk of the row-major route: FMB_RM_CHUNK (32) columns of 2^20 rows, and as many more of a
// shorter transform as make the same 256 MiB (chunks of 32 short columns would be launch bound)
static int64_t rm_chunk(int64_t L = (int64_t)1 << 20) {
    static const long v = env_long("FMB_RM_CHUNK", 32);
    return L >= ((int64_t)1 << 20) ? (int64_t)v : (int64_t)v * (((int64_t)1 << 20) / L);
}

// scratch of an apply: the column-major requirement, or - if larger - what the chunked row-major route needs (a chunk's
// column-major scratch plus its transposed input and output); the layout is not known when the caller asks
int64_t ConvEngine::workspace_bytes(int64_t M, size_t csize) const {
    const int64_t cm = workspace_bytes_cm(M, csize);
    if (shape.npass != 2 || !shape.pow2 || rm_chunk() <= 0 || M < 2) return cm;
    const int64_t tc = std::min<int64_t>(M, rm_chunk(L));
    return std::max(cm, workspace_bytes_cm(tc, csize) + (n_in + n_out) * tc * (int64_t)csize);
}

int64_t ConvEngine::workspace_bytes_cm(int64_t M, size_t csize) const {
    if (shape.npass == 1) return 0;
    const int64_t generic = (int64_t)slab_cols(M, csize) * L * (int64_t)csize;
    if (shape.pow2) {
        int cols, ns;
        slab_plan(M, csize, true, cols, ns);
        int64_t w = std::max(generic, (int64_t)cols * ns * L * (int64_t)csize);
        if (csize == sizeof(float2)) w = std::max(w, v32p_workspace_bytes(M));
        return w;
    }
    return generic;
}

// internal streams of the pipelined-slab schedule (common.h: PipeScope): one set per HOST THREAD and device, shared by all
// plans that thread applies.  Concurrent callers (different host threads, distinct workspaces) therefore issue their
// applies independently - round 1 had one set per device behind a mutex that serialised the host-side issue.
struct StreamPool {
    cudaStream_t s[FMB_MAX_PIPE] = {};
    cudaEvent_t fork = nullptr, join[FMB_MAX_PIPE] = {};
    bool ready = false;
    int ensure() {                       // on the calling thread's pool of the current device
        if (ready) return FMB_OK;
        for (int i = 0; i < FMB_MAX_PIPE; ++i) {
            FMB_CUDA_OK(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
            FMB_CUDA_OK(cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming));
        }
        FMB_CUDA_OK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        ready = true;
        return FMB_OK;
    }
    ~StreamPool() {                      // thread exit; errors (context already gone) are of no consequence
        if (!ready) return;
        for (int i = 0; i < FMB_MAX_PIPE; ++i) { cudaStreamDestroy(s[i]); cudaEventDestroy(join[i]); }
        cudaEventDestroy(fork);
        (void)cudaGetLastError();
    }
};
constexpr int FMB_MAX_DEVICES = 64;
static thread_local StreamPool g_pools[FMB_MAX_DEVICES];     // one per device ordinal (streams belong to a device)

int PipeScope::begin(int ns_, cudaStream_t st) {
    caller = st;
    ns = std::max(1, std::min(ns_, (int)FMB_MAX_PIPE));
    if (ns == 1) return FMB_OK;
    int dev = 0;
    FMB_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= FMB_MAX_DEVICES) { ns = 1; return FMB_OK; }
    StreamPool &pool = g_pools[dev];
    pool_ = &pool;
    int rc = pool.ensure();
    if (rc) return rc;
    FMB_CUDA_OK(cudaEventRecord(pool.fork, st));
    for (int i = 0; i < ns; ++i) FMB_CUDA_OK(cudaStreamWaitEvent(pool.s[i], pool.fork, 0));
    open = true;
    return FMB_OK;
}
cudaStream_t PipeScope::stream(int64_t k) const { return ns > 1 ? static_cast<StreamPool *>(pool_)->s[k % ns] : caller; }
int PipeScope::end() {
    if (ns == 1 || !open) return FMB_OK;
    open = false;
    StreamPool &pool = *static_cast<StreamPool *>(pool_);
    for (int i = 0; i < ns; ++i) {
        FMB_CUDA_OK(cudaEventRecord(pool.join[i], pool.s[i]));
        FMB_CUDA_OK(cudaStreamWaitEvent(caller, pool.join[i], 0));
    }
    return FMB_OK;
}

// ------------------------------------------------------------------------------------------- pass construction
struct TileChoice { int T, NT; size_t smem; int sf, st, psh, pamt; };

static TileChoice choose_tile(const PassGeom &g, size_t csize, bool pow2, bool any_j, bool wide_t, int64_t lines) {
    TileChoice tc;
    const int R = g.R;
    int64_t maxT = std::max<int64_t>(1, pass_limit(g) / R);
    static const long tile_elems = env_long("FMB_TILE_ELEMS", 4096), wide_min = env_long("FMB_WIDE_T", 8);
    int64_t T = std::max<int64_t>(1, tile_elems / R);
    if (wide_t) T = std::max<int64_t>(T, wide_min);
    T = std::min(T, maxT);
    // do not make tiles much wider than the work available
    int64_t lp = pow2 ? next_pow2(lines) : lines;
    T = std::max<int64_t>(1, std::min(T, lp));
    if (pow2) { int64_t p = 1; while (p * 2 <= T) p *= 2; T = p; }
    // shared-memory budget (<= 96 KB per CTA so that at least two CTAs fit an SM)
    while (T > 1 && (size_t)(R + (R >> 4) + 8) * (size_t)T * csize > (size_t)96 << 10) T = pow2 ? T / 2 : T - 1;
    tc.T = (int)T;
    int64_t nt = ((int64_t)R * T + g.min_pnb - 1) / g.min_pnb;
    nt = ((nt + 31) / 32) * 32;
    tc.NT = (int)std::min<int64_t>(std::max<int64_t>(nt, 32), FMB_MAX_NT);
    if (any_j) {            // layout [t][f]: transform index contiguous, one pad element per 16
        tc.sf = 1; tc.psh = 4; tc.pamt = 1;
        tc.st = R + (R >> 4) + 4;
        tc.smem = (size_t)tc.st * (size_t)T * csize;
    } else {                // layout [f][t]: line index contiguous, one pad row per 16 rows when rows are short
        tc.sf = (int)T; tc.st = 1; tc.psh = 4;
        tc.pamt = ((size_t)T * csize < 128) ? (int)T : 0;
        tc.smem = ((size_t)R * T + (size_t)(R >> 4) * tc.pamt + T) * csize;
    }
    return tc;
}

template <typename C>
static int launch_pass(PassParams<C> &p, const PassGeom &g, bool pow2, bool ord_first, bool ord_inner, bool ord_last, cudaStream_t st) {
    const bool any_j = !(ord_first && ord_inner && ord_last);
    // NOTE: the per-stage thread order is carried in t_fastest bits: bit0 first, bit1 inner, bit2 last
    TileChoice tc = choose_tile(g, sizeof(C), pow2, any_j, ord_first || ord_last, p.lines_total);
    p.R = g.R;
    p.T = tc.T;
    p.nstages = (int)g.radix.size();
    for (int s = 0; s < p.nstages; ++s) p.radix[s] = g.radix[s];
    p.t_fastest = (ord_first ? 1 : 0) | (ord_inner ? 2 : 0) | (ord_last ? 4 : 0);
    p.sf = tc.sf; p.st = tc.st; p.psh = tc.psh; p.pamt = tc.pamt;
    long long tiles = (p.lines_total + p.T - 1) / p.T;
    if (tiles <= 0) return FMB_OK;
    if (tiles > 2147483647LL) { set_error("too many tiles"); return FMB_ERR_VALUE; }
#ifdef FMB_EMULATE
    (void)st;
    emulate_launch(tiles, tc.NT, tc.smem, [&](long long tile, int tid, int nt, void *sm, const std::function<void()> &bar) {
        struct HostSync { const std::function<void()> &b; void operator()() const { b(); } } sync{bar};
        if (pow2) pass_body<C, true, HostSync>(p, tile, tid, nt, (C *)sm, sync);
        else pass_body<C, false, HostSync>(p, tile, tid, nt, (C *)sm, sync);
    });
    g_launches.fetch_add(1);
    return FMB_OK;
#else
    return launch_fft_kernel(p, pow2, (unsigned)tiles, tc.NT, tc.smem, st);
#endif
}

template <typename C> static PassParams<C> blank_params() {
    PassParams<C> p;
    memset(&p, 0, sizeof(p));
    p.scale = 1;
    return p;
}

// ------------------------------------------------------------------------------------------- fast path (fft_fast.cuh)
// pass variants (must match fft_fast_inst.cuh)
static const unsigned FV_A_F_ = FO_LOAD_T | FO_TWIDDLE, FV_A_FC_ = FV_A_F_ | FO_IN_CONJ, FV_A_M_ = FV_A_F_ | FO_IN_MASK,
                      FV_A_MP_ = FV_A_M_ | FO_PRE, FV_A_MPC_ = FV_A_MP_ | FO_PRE_CONJ,
                      FV_B_F_ = FO_LOAD_T | FO_STORE_T | FO_OUT_MASK, FV_B_FC_ = FV_B_F_ | FO_OUT_CONJ,
                      FV_BM_ = FO_LOAD_T | FO_STORE_T | FO_TWO_FFTS | FO_TWIDDLE, FV_BMC_ = FV_BM_ | FO_MID_CONJ,
                      FV_C_M_ = FO_STORE_T | FO_OUT_CONJ | FO_OUT_MASK, FV_C_MP_ = FV_C_M_ | FO_POST,
                      FV_C_MPC_ = FV_C_MP_ | FO_POST_CONJ, FV_K_AC_ = FV_B_F_ | FO_IN_CONJ, FV_K_B_ = FO_OUT_MASK,
                      FV_K_BC_ = FO_OUT_MASK | FO_OUT_CONJ, FV_1_FC_ = FO_IN_CONJ | FO_OUT_CONJ | FO_OUT_MASK,
                      FV_1_M_ = FO_TWO_FFTS | FO_IN_MASK | FO_OUT_CONJ | FO_OUT_MASK, FV_1_MC_ = FV_1_M_ | FO_MID_CONJ;
int launch_fast_f32_L6(unsigned opt, const FastArgs<float2> &a, unsigned tiles, cudaStream_t st);
int launch_fast_f64_L6(unsigned opt, const FastArgs<double2> &a, unsigned tiles, cudaStream_t st);
int launch_fast_f32_L7(unsigned opt, const FastArgs<float2> &a, unsigned tiles, cudaStream_t st);
int launch_fast_f64_L7(unsigned opt, const FastArgs<double2> &a, unsigned tiles, cudaStream_t st);
int launch_fast_f32_L8(unsigned opt, const FastArgs<float2> &a, unsigned tiles, cudaStream_t st);
int launch_fast_f32_L9(unsigned opt, const FastArgs<float2> &a, unsigned tiles, cudaStream_t st);
int launch_fast_f32_L10(unsigned opt, const FastArgs<float2> &a, unsigned tiles, cudaStream_t st);
int launch_fast_f32_L11(unsigned opt, const FastArgs<float2> &a, unsigned tiles, cudaStream_t st);
int launch_fast_f32_L12(unsigned opt, const FastArgs<float2> &a, unsigned tiles, cudaStream_t st);
int launch_fast_f64_L8(unsigned opt, const FastArgs<double2> &a, unsigned tiles, cudaStream_t st);
int launch_fast_f64_L9(unsigned opt, const FastArgs<double2> &a, unsigned tiles, cudaStream_t st);
int launch_fast_f64_L10(unsigned opt, const FastArgs<double2> &a, unsigned tiles, cudaStream_t st);
int launch_fast_f64_L11(unsigned opt, const FastArgs<double2> &a, unsigned tiles, cudaStream_t st);

static bool fast_has(const float2 *, int logr) { return logr >= 6 && logr <= 12; }
static bool fast_has(const double2 *, int logr) { return logr >= 6 && logr <= 11; }
static int fast_logt(const float2 *, int logr) { return (logr <= 9) ? (12 - logr) : (FMB_FAST_TILE_LOG2 - logr); }
static int fast_logt(const double2 *, int logr) { return ((logr <= 9) ? (12 - logr) : (FMB_FAST_TILE_LOG2 - logr)) - 1; }
static int fast_launch(int logr, unsigned opt, const FastArgs<float2> &a, unsigned tiles, cudaStream_t st) {
    switch (logr) {
        case 6: return launch_fast_f32_L6(opt, a, tiles, st);
        case 7: return launch_fast_f32_L7(opt, a, tiles, st);
        case 8: return launch_fast_f32_L8(opt, a, tiles, st);
        case 9: return launch_fast_f32_L9(opt, a, tiles, st);
        case 10: return launch_fast_f32_L10(opt, a, tiles, st);
        case 11: return launch_fast_f32_L11(opt, a, tiles, st);
        default: return launch_fast_f32_L12(opt, a, tiles, st);
    }
}
static int fast_launch(int logr, unsigned opt, const FastArgs<double2> &a, unsigned tiles, cudaStream_t st) {
    switch (logr) {
        case 6: return launch_fast_f64_L6(opt, a, tiles, st);
        case 7: return launch_fast_f64_L7(opt, a, tiles, st);
        case 8: return launch_fast_f64_L8(opt, a, tiles, st);
        case 9: return launch_fast_f64_L9(opt, a, tiles, st);
        case 10: return launch_fast_f64_L10(opt, a, tiles, st);
        default: return launch_fast_f64_L11(opt, a, tiles, st);
    }
}
static int ilog2_host(int64_t v) { int l = 0; while (((int64_t)1 << l) < v) ++l; return l; }

template <typename C> bool ConvEngine::fast_ok(int64_t xrs, int64_t yrs, bool in_real) const {
#ifdef FMB_EMULATE
    return false;
#else
    static const long off = env_long("FMB_NO_FAST", 0);
    if (off || !shape.pow2 || shape.npass != 2 || in_real || xrs != 1 || yrs != 1) return false;
    if (L >= ((int64_t)1 << 30)) return false;
    return fast_has((const C *)nullptr, ilog2_host(shape.g[0].R)) && fast_has((const C *)nullptr, ilog2_host(shape.g[1].R));
#endif
}

// Whole transform in ONE kernel (power-of-two L of 128 ... 4096, column-major operands): a line of the specialised pass
// kernel is a column of the operand, a tile is T neighbouring columns.  Fourier is one transform per line; Circulant and
// Toeplitz are FFT -> spectrum -> conj -> FFT -> conj without leaving shared memory, the Toeplitz zero padding being the
// load mask and its cropping the store mask.  Row-major operands (batch contiguous) use the line-fastest thread order:
// the T columns of a tile are the contiguous direction (8 T bytes per row: full 32-byte sectors from T = 4, i.e. up to
// L = 2048).  Handles the first M - M % T columns; returns how many in `done`.
static int launch_v32(unsigned opt, const FastArgs<float2> &a, unsigned lines, cudaStream_t st);
template <typename C>
int ConvEngine::run_single_fast(Dev &d, int direction, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs,
                                int64_t ycs, int64_t M, int64_t &done, cudaStream_t st) const {
    done = 0;
#ifdef FMB_EMULATE
    return FMB_OK;
#else
    static const long off = env_long("FMB_NO_FAST", 0), off1 = env_long("FMB_NO_FAST1", 0), no_v32 = env_long("FMB_NO_V32", 0);
    const int l = ilog2_host(L);
    if (off || off1 || !shape.pow2 || shape.npass != 1 || !fast_has((const C *)nullptr, l)) return FMB_OK;
    const bool chirp = !pre.empty() || !post.empty();                  // chirp-z (Bluestein): pre- and post-multiply around the convolution
    if ((chirp && (pre.empty() || post.empty() || !two_ffts)) || (two_ffts && mid.empty())) return FMB_OK;
    if constexpr (sizeof(C) == sizeof(float2)) if (!chirp) {
        // Convolutions of FFT length 1024 (Circulant(1024); Toeplitz padded to 1024) on column-major batches: the middle
        // pass of the 2^20 transforms IS this operator (FFT -> spectrum -> FFT on contiguous, warp-private lines, 32 values
        // per thread, fused twiddle butterflies) with the same spectrum for every line and the column stride as line stride
        // - 1.8x the 16-value kernel's rate.  Zero padding to exactly twice the length is pruned (never loaded / stored).
        const int64_t rin = direction == FMB_BACKWARD ? n_out : n_in, rout = direction == FMB_BACKWARD ? n_in : n_out;
        if (L == 1024 && two_ffts && !no_v32 && xrs == 1 && yrs == 1 && xcs < ((int64_t)1 << 20) && ycs < ((int64_t)1 << 20) &&
            (M & ~(int64_t)7) > 0 && M < ((int64_t)1 << 31)) {
            const int64_t Mf = M & ~(int64_t)7;
            FastArgs<float2> a;
            memset(&a, 0, sizeof(a));
            a.ncols = (int)((Mf + 1023) / 1024);
            a.in = (const float2 *)x; a.in_cs = 1024 * xcs; a.in_fs = 1; a.in_is = (int)xcs;
            a.out = (float2 *)y; a.out_cs = 1024 * ycs; a.out_ks = 1; a.out_is = (int)ycs;
            a.I = 1024; a.logI = 10;
            a.in_n = (int)rin; a.in_lf = 1; a.in_li = 0;
            a.out_n = (int)rout; a.out_lk = 1; a.out_li = 0;
            a.mid = (const float2 *)d.mid.p; a.mid_is = 0;
            a.tw = (const float2 *)d.twV[0].p;
            const bool bwd1 = direction == FMB_BACKWARD;
            unsigned opt = bwd1 ? V32_1KC : V32_1K;
            if (rin == L && rout == L) opt = bwd1 ? V32_1MC : V32_1M;
            else if (rin * 2 == L && rout * 2 == L) opt = bwd1 ? V32_1HC : V32_1H;
            int rc = launch_v32(opt, a, (unsigned)Mf, st);
            if (rc == FMB_OK) done = Mf;
            return rc;
        }
    }
    const int lt = fast_logt((const C *)nullptr, l);
    const int64_t T = (int64_t)1 << lt, Mf = M & ~(T - 1);
    const bool rm = !(xrs == 1 && yrs == 1);
    if (rm && !(xcs == 1 && ycs == 1)) return FMB_OK;
    const int64_t lim = (int64_t)1 << 30;
    if (Mf == 0 || Mf >= lim || xcs * T >= lim || ycs * T >= lim || xrs >= lim || yrs >= lim) return FMB_OK;
    const bool bwd = direction == FMB_BACKWARD;
    FastArgs<C> a;
    memset(&a, 0, sizeof(a));
    a.ncols = (int)Mf;
    a.in = (const C *)x; a.out = (C *)y;
    if (!rm) {
        a.in_cs = T * xcs; a.in_fs = 1; a.in_is = (int)xcs;
        a.out_cs = T * ycs; a.out_ks = 1; a.out_is = (int)ycs;
        a.I = (int)T; a.logI = lt;
    } else {
        a.in_cs = 0; a.in_fs = (int)xrs; a.in_is = 1;                  // one "column group" holding every line
        a.out_cs = 0; a.out_ks = (int)yrs; a.out_is = 1;
        a.I = (int)lim; a.logI = 30;
    }
    a.in_n = (int)(bwd ? n_out : n_in); a.in_lf = 1; a.in_li = 0;
    a.out_n = (int)(bwd ? n_in : n_out); a.out_lk = 1; a.out_li = 0;
    a.mid = (const C *)d.mid.p; a.mid_is = 0;
    a.tw = (const C *)d.twF[0].p;
    if (chirp) { a.pre = (const C *)(bwd ? d.post.p : d.pre.p); a.post = (const C *)(bwd ? d.pre.p : d.post.p); }   // the adjoint swaps and conjugates them
    const unsigned opt = (two_ffts ? (bwd ? FV_1_MC_ : FV_1_M_) : (bwd ? FV_1_FC_ : FV_K_B_)) | (rm ? (FO_LOAD_T | FO_STORE_T) : 0u) |
                         (chirp ? (FO_PRE | FO_POST | (bwd ? (FO_PRE_CONJ | FO_POST_CONJ) : 0u)) : 0u);
    int rc = fast_launch(l, opt, a, (unsigned)(Mf >> lt), st);
    if (rc == FMB_OK) done = Mf;
    return rc;
#endif
}

// column-major power-of-two transforms: the same three (two) passes as below with compile-time geometry
template <typename C>
int ConvEngine::run_fast(Dev &d, int direction, const void *x, int64_t xcs, void *y, int64_t ycs, int64_t M, void *ws, cudaStream_t st) const {
#ifdef FMB_EMULATE
    return FMB_ERR_NOTIMPL;
#else
    const bool bwd = direction == FMB_BACKWARD;
    const int64_t rows_in = bwd ? n_out : n_in, rows_out = bwd ? n_in : n_out;
    const C *pre_d = (const C *)(bwd ? d.post.p : d.pre.p);
    const C *post_d = (const C *)(bwd ? d.pre.p : d.post.p);
    const int R1 = shape.g[0].R, R2 = shape.g[1].R;
    const int l1 = ilog2_host(R1), l2 = ilog2_host(R2);
    const int t1 = fast_logt((const C *)nullptr, l1), t2 = fast_logt((const C *)nullptr, l2);
    int slab_i, ns;
    slab_plan(M, sizeof(C), true, slab_i, ns);
    const int64_t slab = slab_i;
    int rc;
    // pipelined slabs: fork the caller's stream into ns internal streams, slab k runs on stream k % ns over ring slot
    // k % ns of the workspace, and the caller's stream joins them all at the end (stream-ordered for the caller)
    // plain transforms keep their intermediate in y itself: the second pass reads and writes the same addresses (same
    // geometry on both sides), so there is no ring in the L2 working set (FMB_FAST_INPLACE=0: ring in the workspace)
    static const long inplace_env = env_long("FMB_FAST_INPLACE", 1);
    const bool inplace = inplace_env && !two_ffts && rows_out == L && rows_in == L && (const void *)x != (const void *)y && ycs >= L;
    const int64_t wcs = inplace ? ycs : L;
    PipeScope pipe;
    if ((rc = pipe.begin(ns, st))) return rc;
    void *const ws_base = ws;
    int64_t slab_idx = 0;
    for (int64_t c0 = 0; c0 < M; c0 += slab, ++slab_idx) {
        const int64_t nc = std::min(slab, M - c0);
        if (ns > 1) {
            st = pipe.stream(slab_idx);
            ws = (char *)ws_base + (size_t)(slab_idx % ns) * (size_t)slab * (size_t)L * sizeof(C);
        }
        if (inplace) ws = (C *)y + c0 * ycs;
        FastArgs<C> base;
        memset(&base, 0, sizeof(base));
        base.ncols = (int)nc;
        base.twL = (const C *)d.twL.p; base.twH = (const C *)d.twH.p; base.tw_shift = d.tw_shift;
        base.tw_mask = (unsigned)(((int64_t)1 << d.tw_shift) - 1);
        if (kron_a > 0) {
            // Kron(Fourier(R1), Fourier(R2)): 2-D transform of the row-major (R1, R2) image, no twiddle between the passes
            FastArgs<C> a = base;                                     // over i1 (stride R2), lines i2; natural order out
            a.in = (const C *)x + c0 * xcs; a.in_cs = xcs; a.in_fs = R2; a.in_is = 1;
            a.out = (C *)ws; a.out_cs = wcs; a.out_ks = R2; a.out_is = 1;
            a.I = R2; a.logI = l2;
            a.out_n = (int)L; a.out_lk = R2; a.out_li = 1;
            a.tw = (const C *)d.twF[0].p;
            if ((rc = fast_launch(l1, bwd ? FV_K_AC_ : FV_B_F_, a, (unsigned)((nc * R2) >> t1), st))) return rc;
            FastArgs<C> b2 = base;                                    // over i2 (contiguous), lines k1
            b2.in = (const C *)ws; b2.in_cs = wcs; b2.in_fs = 1; b2.in_is = R2;
            b2.out = (C *)y + c0 * ycs; b2.out_cs = ycs; b2.out_ks = 1; b2.out_is = R2;
            b2.I = R1; b2.logI = l1;
            b2.out_n = (int)L; b2.out_lk = 1; b2.out_li = R2;
            b2.tw = (const C *)d.twF[1].p;
            if ((rc = fast_launch(l2, bwd ? FV_K_BC_ : FV_K_B_, b2, (unsigned)((nc * R1) >> t2), st))) return rc;
            continue;
        }
        // ---- pass A: length R1 over n = f*R2 + i, lines i < R2; out: ws[c][i][k] (k contiguous)
        {
            FastArgs<C> a = base;
            a.in = (const C *)x + c0 * xcs; a.in_cs = xcs; a.in_fs = R2; a.in_is = 1;
            a.out = (C *)ws; a.out_cs = wcs; a.out_ks = 1; a.out_is = R1;
            a.I = R2; a.logI = l2;
            a.in_n = (int)rows_in; a.in_lf = R2; a.in_li = 1;
            a.tw = (const C *)d.twF[0].p; a.twS = (const C *)d.twS[0].p;
            a.pre = pre_d;
            unsigned opt;
            if (!two_ffts) opt = bwd ? FV_A_FC_ : FV_A_F_;
            else if (pre_d) opt = bwd ? FV_A_MPC_ : FV_A_MP_;
            else opt = (rows_in < L) ? FV_A_M_ : FV_A_F_;
            if ((rc = fast_launch(l1, opt, a, (unsigned)((nc * R2) >> t1), st))) return rc;
        }
        if (!two_ffts) {
            // ---- pass B: length R2 over f = n2 (stride R1 in ws), lines i = k1 < R1; out y[k*R1 + i]
            FastArgs<C> a = base;
            a.in = (const C *)ws; a.in_cs = wcs; a.in_fs = R1; a.in_is = 1;
            a.out = (C *)y + c0 * ycs; a.out_cs = ycs; a.out_ks = R1; a.out_is = 1;
            a.I = R1; a.logI = l1;
            a.out_n = (int)rows_out; a.out_lk = R1; a.out_li = 1;
            a.tw = (const C *)d.twF[1].p;
            if ((rc = fast_launch(l2, bwd ? FV_B_FC_ : FV_B_F_, a, (unsigned)((nc * R1) >> t2), st))) return rc;
        } else {
            {   // ---- pass B': in place on ws; FFT over n2, * spectrum[k*R1 + i], conj, FFT, * W^{ik}
                FastArgs<C> a = base;
                a.in = (const C *)ws; a.in_cs = L; a.in_fs = R1; a.in_is = 1;
                a.out = (C *)ws; a.out_cs = L; a.out_ks = R1; a.out_is = 1;
                a.I = R1; a.logI = l1;
                a.mid = (const C *)d.mid.p; a.mid_is = R2;
                a.tw = (const C *)d.twF[1].p; a.twS = (const C *)d.twS[1].p;
                if ((rc = fast_launch(l2, bwd ? FV_BMC_ : FV_BM_, a, (unsigned)((nc * R1) >> t2), st))) return rc;
            }
            {   // ---- pass C: length R1 over f = k1 (contiguous in ws rows), lines i = m2 < R2; out y[k*R2 + i]
                FastArgs<C> a = base;
                a.in = (const C *)ws; a.in_cs = L; a.in_fs = 1; a.in_is = R1;
                a.out = (C *)y + c0 * ycs; a.out_cs = ycs; a.out_ks = R2; a.out_is = 1;
                a.I = R2; a.logI = l2;
                a.out_n = (int)rows_out; a.out_lk = R2; a.out_li = 1;
                a.post = post_d;
                a.tw = (const C *)d.twF[0].p;
                unsigned opt = post_d ? (bwd ? FV_C_MPC_ : FV_C_MP_) : FV_C_M_;
                if ((rc = fast_launch(l1, opt, a, (unsigned)((nc * R2) >> t1), st))) return rc;
            }
        }
    }
    if ((rc = pipe.end())) return rc;
    return FMB_OK;
#endif
}

// ------------------------------------------------------------------------------------------- V32 path (fft_v32.cuh)
int launch_v32_a(unsigned opt, const FastArgs<float2> &a, unsigned lines, int shape, cudaStream_t st);
int launch_v32_b(unsigned opt, const FastArgs<float2> &a, unsigned lines, int shape, cudaStream_t st);
int launch_v32_m(unsigned opt, const FastArgs<float2> &a, unsigned lines, int shape, cudaStream_t st);
int launch_v32_c(unsigned opt, const FastArgs<float2> &a, unsigned lines, int shape, cudaStream_t st);
int launch_v32_1(unsigned opt, const FastArgs<float2> &a, unsigned lines, int shape, cudaStream_t st);
// tile / occupancy instantiation (fft_v32.cuh: launch_v32_variant): FMB_V32_OCC for the strided passes, FMB_V32_MSHAPE for
// the middle pass of a convolution
static int launch_v32(unsigned opt, const FastArgs<float2> &a, unsigned lines, cudaStream_t st) {
    static const long occ = env_long("FMB_V32_OCC", 0), mshape = env_long("FMB_V32_MSHAPE", 0);
    const int shape = (int)((opt & FO_TWO_FFTS) ? mshape : occ);
    int rc = launch_v32_a(opt, a, lines, shape, st);
    if (rc == FMB_ERR_NOTIMPL) rc = launch_v32_b(opt, a, lines, shape, st);
    if (rc == FMB_ERR_NOTIMPL) rc = launch_v32_m(opt, a, lines, shape, st);
    if (rc == FMB_ERR_NOTIMPL) rc = launch_v32_c(opt, a, lines, shape, st);
    if (rc == FMB_ERR_NOTIMPL) rc = launch_v32_1(opt, a, lines, shape, st);
    if (rc == FMB_ERR_NOTIMPL) set_error("V32 path: unknown pass variant %u", opt);
    return rc;
}

int launch_v32t(unsigned opt, const FastArgs<float2> &a, const CUtensorMap &map, int map_col0, unsigned lines, cudaStream_t st);
static int v32p_tensor_map(CUtensorMap *map, const void *base, int64_t rows, int64_t col_stride, int64_t cols);
static bool v32p_tma_available();

bool ConvEngine::v32_ok(size_t csize) const {
#ifdef FMB_EMULATE
    return false;
#else
    static const long off = env_long("FMB_NO_V32", 0);
    return !off && csize == sizeof(float2) && shape.npass == 2 && shape.g[0].R == 1024 && shape.g[1].R == 1024;
#endif
}

// L = 2^20 = 1024 x 1024, complex64, column-major: the same passes as run_fast with 32 values per thread.  The
// intermediate is stored [k1][n2] (n2 contiguous), so that the middle pass of a convolution works on contiguous lines.
int ConvEngine::run_v32(Dev &d, int direction, const void *x, int64_t xcs, void *y, int64_t ycs, int64_t M, void *ws, cudaStream_t st) const {
#ifdef FMB_EMULATE
    return FMB_ERR_NOTIMPL;
#else
    typedef float2 C;
    const bool bwd = direction == FMB_BACKWARD;
    const int64_t rows_in = bwd ? n_out : n_in, rows_out = bwd ? n_in : n_out;
    const C *pre_d = (const C *)(bwd ? d.post.p : d.pre.p);
    const C *post_d = (const C *)(bwd ? d.pre.p : d.post.p);
    const int R1 = 1024, R2 = 1024, l1 = 10, l2 = 10;
    // the four-step twiddle between the inverse transform's two passes rides on the LAST pass's loads (FMB_V32_TWM=1: on
    // the middle pass's stores, as in round 1): the middle pass is the FP32-bound one, the last pass waits for its loads
    static const bool tw_in_c = env_long("FMB_V32_TWM", 0) == 0;
    int slab_i, ns;
    slab_plan(M, sizeof(C), true, slab_i, ns);
    const int64_t slab = slab_i;
    int rc;
    // FMB_V32T: the strided passes of a convolution (first: x -> ring, last: ring -> y) fetch their tiles with the TMA engine
    // (fft_v32p.cuh: v32t_pass_kernel) instead of 32 register-direct loads per thread; 1: both, 2: first only, 3: last only
    static const long v32t = env_long("FMB_V32T", 2);       // as measured (profiles/r2_experiments.txt, call 15): first pass only
    CUtensorMap map_x, map_ring;
    bool tma_a = false, tma_c = false;
    static const long v32t_plain = env_long("FMB_V32T_PLAIN", 0);       // experiments: also the first pass of a plain transform
    if (v32t > 0 && (two_ffts || v32t_plain) && kron_a == 0 && !pre_d && !post_d && v32p_tma_available() && rows_in > 0 && rows_in % 1024 == 0 &&
        !(reinterpret_cast<uintptr_t>(x) & 15) && !(xcs & 1) && xcs >= rows_in && !(reinterpret_cast<uintptr_t>(ws) & 127)) {
        // (a zero-padded input - Toeplitz - is better served by the pruned register-direct pass unless asked for with 1)
        tma_a = (v32t == 1 || (v32t == 2 && rows_in == L)) && v32p_tensor_map(&map_x, x, rows_in, xcs, M) == FMB_OK;
        tma_c = two_ffts && (v32t == 1 || v32t == 3) && v32p_tensor_map(&map_ring, ws, L, L, (int64_t)std::max(ns, 1) * slab) == FMB_OK;
    }
    // Convolutions that keep all L rows (Circulant) run IN PLACE on y: pass A writes y, the middle pass transforms its
    // lines there and the last pass reads and writes the same addresses (same strided geometry on both sides).  No ring in
    // the L2 working set, and no dirty ring lines for the L2 to write back to DRAM (FMB_V32_INPLACE=0: ring).
    static const long inplace_env = env_long("FMB_V32_INPLACE", 1);
    const bool inplace = inplace_env && two_ffts && kron_a == 0 && rows_out == L && (const void *)x != (const void *)y && ycs >= L;
    const int64_t wcs = inplace ? ycs : L;                        // column stride of the intermediate
    if (inplace) tma_c = false;
    PipeScope pipe;
    if ((rc = pipe.begin(ns, st))) return rc;
    // FMB_L2_PERSIST=1 (experiments): the ring of intermediates is the only data with reuse - ask the L2 to keep it
    // (persisting access-policy window on the internal streams; everything else those kernels touch is streaming)
    static const long l2_persist = env_long("FMB_L2_PERSIST", 0);
    if (l2_persist > 0 && ns > 1) {
        static bool limit_set = false;
        if (!limit_set) {
            int max_persist = 0, dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
            limit_set = true;
        }
        int max_window = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.base_ptr = ws;
        attr.accessPolicyWindow.num_bytes = std::min<size_t>((size_t)ns * (size_t)slab * (size_t)L * sizeof(C), (size_t)max_window);
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = l2_persist == 2 ? cudaAccessPropertyNormal : cudaAccessPropertyStreaming;
        for (int i = 0; i < ns; ++i) cudaStreamSetAttribute(pipe.stream(i), cudaStreamAttributeAccessPolicyWindow, &attr);
    }
    void *const ws_base = ws;
    int64_t slab_idx = 0;
    for (int64_t c0 = 0; c0 < M; c0 += slab, ++slab_idx) {
        const int64_t nc = std::min(slab, M - c0);
        int ring_col0 = 0;
        if (ns > 1) {
            st = pipe.stream(slab_idx);
            ws = (char *)ws_base + (size_t)(slab_idx % ns) * (size_t)slab * (size_t)L * sizeof(C);
            ring_col0 = (int)((slab_idx % ns) * slab);
        }
        if (inplace) ws = (C *)y + c0 * ycs;
        const unsigned tiles = (unsigned)(nc * 1024);          // lines; the launcher divides by its tile width
        FastArgs<C> base;
        memset(&base, 0, sizeof(base));
        base.ncols = (int)nc;
        base.twL = (const C *)d.twL.p; base.twH = (const C *)d.twH.p; base.tw_shift = d.tw_shift;
        base.tw_mask = (unsigned)(((int64_t)1 << d.tw_shift) - 1);
        base.I = 1024; base.logI = 10;
        if (kron_a > 0) {
            FastArgs<C> a = base;                                     // over i1 (stride R2), lines i2; natural order out
            a.in = (const C *)x + c0 * xcs; a.in_cs = xcs; a.in_fs = R2; a.in_is = 1;
            a.out = (C *)ws; a.out_cs = L; a.out_ks = R2; a.out_is = 1;
            a.out_n = (int)L; a.out_lk = R2; a.out_li = 1;
            a.tw = (const C *)d.twV[0].p;
            if ((rc = launch_v32(bwd ? V32_K_AC : V32_K_A, a, tiles, st))) return rc;
            FastArgs<C> b2 = base;                                    // over i2 (contiguous), lines k1
            b2.in = (const C *)ws; b2.in_cs = L; b2.in_fs = 1; b2.in_is = R2;
            b2.out = (C *)y + c0 * ycs; b2.out_cs = ycs; b2.out_ks = 1; b2.out_is = R2;
            b2.out_n = (int)L; b2.out_lk = 1; b2.out_li = R2;
            b2.tw = (const C *)d.twV[1].p;
            if ((rc = launch_v32(bwd ? V32_K_BC : V32_K_B, b2, tiles, st))) return rc;
            continue;
        }
        {   // ---- pass A: length R1 over n = f*R2 + i (lines i contiguous); out ws[k1*R2 + i], times W^{i k1}
            FastArgs<C> a = base;
            a.in = (const C *)x + c0 * xcs; a.in_cs = xcs; a.in_fs = R2; a.in_is = 1;
            a.out = (C *)ws; a.out_cs = wcs; a.out_ks = R2; a.out_is = 1;
            a.in_n = (int)rows_in; a.in_lf = R2; a.in_li = 1;
            a.tw = (const C *)d.twV[0].p; a.twS = (const C *)d.twS32[0].p;
            a.pre = pre_d;
            // L2 prefetch of the input of the slab `pf_ahead` slabs later (full slabs of full columns only)
            static const long pf_ahead = env_long("FMB_V32_PF", 0);
            if (pf_ahead > 0 && rows_in == L && c0 + (pf_ahead + 1) * slab <= M && nc == slab)
                a.pf = (const C *)x + (c0 + pf_ahead * slab) * xcs;
            unsigned opt;
            if (!two_ffts) opt = bwd ? V32_A_FC : V32_A_F;
            else if (pre_d) opt = bwd ? V32_A_MPC : V32_A_MP;
            else opt = (rows_in < L) ? V32_A_M : V32_A_F;
            // zero padding to exactly twice the length (Toeplitz n = m = L/2): the padded half is never loaded and the first
            // radix-2 level of the butterflies is skipped (FMB_V32_PRUNE=0: masked loads, full butterflies)
            static const bool prune = env_long("FMB_V32_PRUNE", 1) != 0;
            if (prune && two_ffts && !pre_d && rows_in * 2 == L && !tma_a) opt = V32_A_H;
            if (tma_a) {                           // zero padding = rows outside the tensor map
                if ((rc = launch_v32t((bwd && !two_ffts) ? V32_A_FC : V32_A_F, a, map_x, (int)c0, tiles, st))) return rc;
            } else if ((rc = launch_v32(opt, a, tiles, st))) return rc;
        }
        if (!two_ffts) {
            // ---- pass B: length R2 over n2 (contiguous in ws), lines k1; out y[k1 + R1 k2]
            FastArgs<C> a = base;
            a.in = (const C *)ws; a.in_cs = L; a.in_fs = 1; a.in_is = R2;
            a.out = (C *)y + c0 * ycs; a.out_cs = ycs; a.out_ks = R1; a.out_is = 1;
            a.out_n = (int)rows_out; a.out_lk = R1; a.out_li = 1;
            a.tw = (const C *)d.twV[1].p;
            const bool all_rows = rows_out == L;
            if ((rc = launch_v32(bwd ? (all_rows ? V32_B_NC : V32_B_FC) : (all_rows ? V32_B_N : V32_B_F), a, tiles, st))) return rc;
        } else {
            {   // ---- pass B': in place on ws lines k1: FFT over n2, * spectrum[k1][k2], conj, FFT, * W^{k1 m2}
                FastArgs<C> a = base;
                a.in = (const C *)ws; a.in_cs = wcs; a.in_fs = 1; a.in_is = R2;
                a.out = (C *)ws; a.out_cs = wcs; a.out_ks = 1; a.out_is = R2;
                a.mid = (const C *)d.mid.p; a.mid_is = R2;
                a.tw = (const C *)d.twV[1].p; a.twS = (const C *)d.twS32[1].p;
                if ((rc = launch_v32(tw_in_c ? (bwd ? V32_BMC_N : V32_BM_N) : (bwd ? V32_BMC : V32_BM), a, tiles, st))) return rc;
            }
            {   // ---- pass C: length R1 over k1 (stride R2 in ws), lines m2; conj, post-multiply, truncate; out y[m1*R2 + m2]
                FastArgs<C> a = base;
                a.in = (const C *)ws; a.in_cs = wcs; a.in_fs = R2; a.in_is = 1;
                a.out = (C *)y + c0 * ycs; a.out_cs = ycs; a.out_ks = R2; a.out_is = 1;
                a.out_n = (int)rows_out; a.out_lk = R2; a.out_li = 1;
                a.post = post_d;
                a.tw = (const C *)d.twV[0].p; a.twS = (const C *)d.twS32[1].p;
                unsigned opt = post_d ? (bwd ? V32_C_MPC : V32_C_MP) : (rows_out == L ? V32_C_N : V32_C_M);
                static const bool prune = env_long("FMB_V32_PRUNE", 1) != 0;
                if (prune && !post_d && rows_out * 2 == L && !tma_c) opt = V32_C_H;      // the dropped half of the rows is never computed
                if (tw_in_c) opt |= V32_C_TW;
                if (tma_c) {
                    if ((rc = launch_v32t(opt, a, map_ring, ring_col0, tiles, st))) return rc;
                } else if ((rc = launch_v32(opt, a, tiles, st))) return rc;
            }
        }
    }
    return pipe.end();
    (void)l1; (void)l2;
#endif
}


// ------------------------------------------------------------------------------------------- V32P path (fft_v32p.cuh)
int launch_v32p_f(int variant, const V32PArgs &g, const CUtensorMap &mx, const CUtensorMap &mr, cudaStream_t st);
int launch_v32p_c0(int variant, const V32PArgs &g, const CUtensorMap &mx, const CUtensorMap &mr, cudaStream_t st);
int launch_v32p_c1(int variant, const V32PArgs &g, const CUtensorMap &mx, const CUtensorMap &mr, cudaStream_t st);

struct V32PGeom { int slab_cols, delay, nslot, npass, mix; int64_t nslabs, counter_bytes, ring_bytes; unsigned tiles, items_per_step; };
static V32PGeom v32p_geom(int64_t M, int64_t L, bool two) {
    static const long slab_env = env_long("FMB_V32P_SLAB", 0), delay_env = env_long("FMB_V32P_DELAY", 0),
                      mix_env = env_long("FMB_V32P_MIX", 0);
    V32PGeom f;
    f.npass = two ? 3 : 2;
    f.mix = mix_env ? 1 : 0;
    f.slab_cols = slab_env > 0 ? (int)std::min<long>(slab_env, 64) : (two ? 2 : 1);      // as measured (tools/check_v32p.py)
    f.tiles = (unsigned)f.slab_cols * 128u;
    f.items_per_step = (unsigned)f.npass * f.tiles;
    // Correctness does not depend on D (fft_v32p.cuh); speed does: a tile is requested up to three items per CTA ahead of
    // its arithmetic and completions are published a little late, so producers should be >= AHEAD * G items back.
    //   blocked order:     distance = D S + 1;   interleaved order:  distance = (D - 1) S + npass + 1     (S = items per step)
    static const long ahead_env = env_long("FMB_V32P_AHEAD", 5);
    const int64_t G = device_props().sm_count, S = f.items_per_step, need = std::max<long>(1, ahead_env) * G;
    int dmin = 1;
    while ((f.mix ? (int64_t)(dmin - 1) * S + f.npass + 1 : (int64_t)dmin * S + 1) < need) ++dmin;
    f.delay = (int)std::max<long>(dmin, delay_env);
    // the slot of slab s is reused by slab s + nslot, whose pass A waits for the last pass of slab s: same slack again
    f.nslot = (f.npass - 1) * f.delay + 1 + (int)((need + S - 1) / S);
    f.nslabs = (M + f.slab_cols - 1) / f.slab_cols;
    f.counter_bytes = ((int64_t)4 * 3 * f.nslabs + 1023) / 1024 * 1024;
    f.ring_bytes = (int64_t)f.nslot * f.slab_cols * L * (int64_t)sizeof(float2);
    return f;
}

int64_t ConvEngine::v32p_workspace_bytes(int64_t M) const {
    if (!v32_ok(sizeof(float2))) return 0;
    const V32PGeom f = v32p_geom(M, L, two_ffts);
    return f.counter_bytes + f.ring_bytes;
}

typedef CUresult (*fmb_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static fmb_encode_tiled_fn encode_tiled_fn() {
#ifdef FMB_EMULATE
    return nullptr;
#else
    static fmb_encode_tiled_fn fn = []() -> fmb_encode_tiled_fn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (fmb_encode_tiled_fn)p;
    }();
    return fn;
#endif
}

static bool v32p_tma_available() { return encode_tiled_fn() != nullptr; }

// tensor of complex64 (as 8-byte words) [cols][rows / 1024][1024]; a request is a box of 8 words x 256 rows x 1 column
static int v32p_tensor_map(CUtensorMap *map, const void *base, int64_t rows, int64_t col_stride, int64_t cols) {
    static const long promo = env_long("FMB_V32P_PROMO", 0);
    fmb_encode_tiled_fn enc = encode_tiled_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return FMB_ERR_CUDA; }
    const cuuint64_t dims[3] = {1024, (cuuint64_t)(rows / 1024), (cuuint64_t)cols};
    const cuuint64_t strides[2] = {8192, (cuuint64_t)col_stride * sizeof(float2)};
    const cuuint32_t box[3] = {8, 256, 1}, estr[3] = {1, 1, 1};
    const CUtensorMapL2promotion pr = promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                    : promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return FMB_ERR_CUDA; }
    return FMB_OK;
}

bool ConvEngine::v32p_ok(int direction, const void *x, int64_t xcs, const void *y, int64_t ycs) const {
#ifdef FMB_EMULATE
    return false;
#else
    // 1 (default): plain transforms (1-D four-step and Kron's 2-D: measured +16..19 % over the per-pass kernels);
    // 2: convolutions too - correct (bit-identical, tests/test_gpu_parity.py) but not faster so far: the middle pass is
    // bound by arithmetic and shared memory, gains nothing from the asynchronous fetch of its (contiguous) lines and
    // pays one more shared-memory read for it
    static const long on = env_long("FMB_V32P", 1);
    if (!on || !pre.empty() || !post.empty()) return false;
    if (two_ffts && on < 2) return false;
    const int64_t rows_in = direction == FMB_BACKWARD ? n_out : n_in, rows_out = direction == FMB_BACKWARD ? n_in : n_out;
    if (rows_in <= 0 || rows_in % 1024 != 0) return false;              // zero padding = out-of-bounds rows of the tensor copy
    if (!two_ffts && rows_out != L) return false;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (xcs & 1) || xcs < rows_in) return false;
    return encode_tiled_fn() != nullptr;
    (void)y; (void)ycs;
#endif
}

// L = 2^20 complex64, column-major, no chirp: every pass in ONE persistent launch (fft_v32p.cuh)
int ConvEngine::run_v32p(Dev &d, int direction, const void *x, int64_t xcs, void *y, int64_t ycs, int64_t M, void *ws, int64_t ws_bytes,
                         cudaStream_t st) const {
#ifdef FMB_EMULATE
    return FMB_ERR_NOTIMPL;
#else
    typedef float2 C;
    static const long hint_on = env_long("FMB_V32P_HINTS", 1);
    const bool bwd = direction == FMB_BACKWARD;
    const int64_t rows_in = bwd ? n_out : n_in, rows_out = bwd ? n_in : n_out;
    const int R1 = 1024, R2 = 1024;
    const V32PGeom f = v32p_geom(M, L, two_ffts);
    if (ws == nullptr || ws_bytes < f.counter_bytes + f.ring_bytes) {
        set_error("workspace too small: need %lld bytes", (long long)(f.counter_bytes + f.ring_bytes));
        return FMB_ERR_WORKSPACE;
    }
    FMB_CUDA_OK(cudaMemsetAsync(ws, 0, (size_t)f.counter_bytes, st));
    V32PArgs g;
    memset(&g, 0, sizeof(g));
    g.npass = f.npass; g.ncols = (int)M; g.slab_cols = f.slab_cols; g.nslabs = (int)f.nslabs;
    g.delay = f.delay; g.nslot = f.nslot; g.mix = f.mix;
    g.tiles = f.tiles; g.items_per_step = f.items_per_step;
    g.total_items = (unsigned)((f.nslabs + (int64_t)(f.npass - 1) * f.delay) * f.items_per_step);
    g.slot_stride = (long long)f.slab_cols * L;
    g.ring_cs = L;
    g.y_slab_stride = (long long)f.slab_cols * ycs;
    g.L = L;
    g.done = (unsigned *)ws;
    g.ring = (C *)((char *)ws + f.counter_bytes);
    // plain transforms in place: the intermediate is written into y and overwritten there by the second pass - no ring in
    // the L2 working set and no dirty ring lines to write back (FMB_V32P_INPLACE=0: ring)
    static const long inplace_env = env_long("FMB_V32P_INPLACE", 1);
    const bool inplace = inplace_env && (!two_ffts || rows_out == L) && (const void *)x != (const void *)y && (ycs & 1) == 0 &&
                         (reinterpret_cast<uintptr_t>(y) & 15) == 0 && ycs >= L;
    if (inplace) {
        g.ring = (C *)y; g.ring_cs = ycs; g.slot_stride = (long long)f.slab_cols * ycs;
        g.nslot = (int)f.nslabs + 1;                                 // slot = slab: nothing to wait for before pass A
    }
    // L2 policies (createpolicy encodings): x is read once -> evict first; the ring is the working set -> evict last
    g.hint_x = hint_on ? 0x12F0000000000000ull : 0x1000000000000000ull;
    g.hint_ring = hint_on ? 0x14F0000000000000ull : 0x1000000000000000ull;
    FastArgs<C> base;
    memset(&base, 0, sizeof(base));
    base.twL = (const C *)d.twL.p; base.twH = (const C *)d.twH.p; base.tw_shift = d.tw_shift;
    base.tw_mask = (unsigned)(((int64_t)1 << d.tw_shift) - 1);
    base.I = 1024; base.logI = 10;
    {   // pass A: length R1 over n = f*R2 + i; ring[k1*R2 + i] * W^{i k1}    (in place: y[i*R1 + k1])
        FastArgs<C> a = base;
        a.out_cs = L; a.out_ks = R2; a.out_is = 1;
        if (inplace && kron_a == 0 && !two_ffts) { a.out_ks = 1; a.out_is = R1; }
        a.tw = (const C *)d.twV[0].p; a.twS = (const C *)d.twS32[0].p;
        g.pass[0] = a;
    }
    int variant;
    if (kron_a > 0) {
        FastArgs<C> a = base;                       // 2-D transform: second pass over i2 (contiguous), natural order out
        a.out = (C *)y; a.out_cs = ycs; a.out_ks = 1; a.out_is = R2;
        a.out_n = (int)L; a.out_lk = 1; a.out_li = R2;
        a.tw = (const C *)d.twV[1].p;
        g.pass[1] = a;
        variant = bwd ? VP_KC : VP_K;
    } else if (!two_ffts) {
        FastArgs<C> a = base;                       // pass B: lines k1 contiguous in the ring; y[k1 + R1 k2]
        a.out = (C *)y; a.out_cs = ycs; a.out_ks = R1; a.out_is = 1;
        a.out_n = (int)rows_out; a.out_lk = R1; a.out_li = 1;
        a.tw = (const C *)d.twV[1].p;
        g.pass[1] = a;
        variant = inplace ? (bwd ? VP_FIC : VP_FI) : (bwd ? VP_FC : VP_F);
    } else {
        FastArgs<C> a = base;                       // pass B': in place on the lines k1 of the ring
        a.out_cs = L; a.out_ks = 1; a.out_is = R2;
        a.mid = (const C *)d.mid.p; a.mid_is = R2;
        a.tw = (const C *)d.twV[1].p; a.twS = (const C *)d.twS32[1].p;
        g.pass[1] = a;
        FastArgs<C> c = base;                       // pass C: length R1 over k1, lines m2; y[m1*R2 + m2]
        c.out = (C *)y; c.out_cs = ycs; c.out_ks = R2; c.out_is = 1;
        c.out_n = (int)rows_out; c.out_lk = R2; c.out_li = 1;
        c.tw = (const C *)d.twV[0].p;
        g.pass[2] = c;
        variant = (bwd ? VP_CVC_N : VP_CV_N) + (rows_out == L ? 0 : 1);
    }
    CUtensorMap mx, mr;
    int rc;
    if ((rc = v32p_tensor_map(&mx, x, rows_in, xcs, M))) return rc;
    if (inplace) rc = v32p_tensor_map(&mr, y, L, ycs, M);
    else rc = v32p_tensor_map(&mr, g.ring, L, L, (int64_t)f.nslot * f.slab_cols);
    if (rc) return rc;
    rc = launch_v32p_f(variant, g, mx, mr, st);
    if (rc == FMB_ERR_NOTIMPL) rc = launch_v32p_c0(variant, g, mx, mr, st);
    if (rc == FMB_ERR_NOTIMPL) rc = launch_v32p_c1(variant, g, mx, mr, st);
    if (rc == FMB_ERR_NOTIMPL) set_error("V32P path: unknown variant %d", variant);
    return rc;
#endif
}


template <typename C>
int ConvEngine::run_t(Dev &d, int direction, const void *x, int64_t xrs, int64_t xcs, bool in_real, void *y, int64_t yrs,
                      int64_t ycs, int64_t M, void *ws, int64_t ws_bytes, cudaStream_t st) const {
    int rc = ensure_dev<C>(d);
    if (rc) return rc;
    if (M <= 0) return FMB_OK;
    const bool bwd = direction == FMB_BACKWARD;
    const int64_t rows_in = bwd ? n_out : n_in, rows_out = bwd ? n_in : n_out;
    const C *pre_d = (const C *)(bwd ? d.post.p : d.pre.p);      // adjoint swaps and conjugates pre / post
    const C *post_d = (const C *)(bwd ? d.pre.p : d.post.p);
    const C *mid_d = (const C *)d.mid.p;
    const bool rowmajor = (xcs == 1 && M > 1);                   // batch contiguous (torch default) vs fastmat column-major
    const bool pow2 = shape.pow2;

    if (shape.npass == 1) {
        if (!in_real) {
            // specialised single-kernel route for whole tiles of columns; a ragged tail takes the generic kernel below
            int64_t done = 0;
            if ((rc = run_single_fast<C>(d, direction, x, xrs, xcs, y, yrs, ycs, M, done, st))) return rc;
            if (done == M) return FMB_OK;
            x = (const C *)x + done * xcs; y = (C *)y + done * ycs; M -= done;
        }
        PassParams<C> p = blank_params<C>();
        p.two_ffts = two_ffts;
        p.lines_total = M; p.I = 1; p.ncols = M; p.line_c_fastest = 0;
        p.in = x; p.in_real = in_real; p.in_cs = xcs; p.in_rs = xrs; p.in_lf = 1; p.in_li = 0; p.in_n = rows_in;
        p.pre = pre_d; p.pre_conj = bwd;
        p.mid = mid_d; p.mid_conj = bwd; p.mid_lk = 1; p.mid_li = 0;
        p.out = y; p.out_cs = ycs; p.out_rs = yrs; p.out_lk = 1; p.out_li = 0; p.out_n = rows_out;
        p.post = post_d; p.post_conj = bwd;
        if (two_ffts) p.conj_a = 1;
        else if (bwd) { p.in_conj = 1; p.conj_a = 1; }
        p.wR = (const C *)d.wR[0].p;
        return launch_pass<C>(p, shape.g[0], pow2, rowmajor, rowmajor, rowmajor, st);
    }

    // ---------------- row-major (batch-contiguous, torch's default) operands of the power-of-two two-pass shapes: chunks of
    // FMB_RM_CHUNK columns are transposed into the column-major layout, pushed through the fast path and transposed back
    // (the generic strided kernel below needs 5-9x the time of the fast path; two extra sweeps cost ~0.7x).
    if (rowmajor && ycs == 1 && !in_real && fast_ok<C>(1, 1, false) && rm_chunk() > 0) {
        const int64_t tc = std::min<int64_t>(M, rm_chunk(L));
        const int64_t inner = workspace_bytes_cm(tc, sizeof(C));
        const int64_t need = inner + (rows_in + rows_out) * tc * (int64_t)sizeof(C);
        if (ws == nullptr || ws_bytes < need) { set_error("workspace too small: need %lld bytes", (long long)need); return FMB_ERR_WORKSPACE; }
        C *xt = (C *)((char *)ws + inner), *yt = xt + rows_in * tc;
        for (int64_t c0 = 0; c0 < M; c0 += tc) {
            const int64_t nc = std::min(tc, M - c0);
            if ((rc = transpose_apply((const C *)x + c0, xrs, xt, rows_in, rows_in, nc, sizeof(C), st))) return rc;
            if ((rc = run_t<C>(d, direction, xt, 1, rows_in, false, yt, 1, rows_out, nc, ws, inner, st))) return rc;
            if ((rc = transpose_apply(yt, rows_out, (C *)y + c0, yrs, nc, rows_out, sizeof(C), st))) return rc;
        }
        return FMB_OK;
    }

    // ---------------- two shared-memory passes per transform, slab by slab over an L2-resident intermediate
    const int64_t slab = slab_cols(M, sizeof(C));
    if (fast_ok<C>(xrs, yrs, in_real)) {
        if (ws == nullptr || ws_bytes < workspace_bytes_cm(M, sizeof(C))) {
            set_error("workspace too small: need %lld bytes", (long long)workspace_bytes_cm(M, sizeof(C)));
            return FMB_ERR_WORKSPACE;
        }
        if (v32_ok(sizeof(C)) && v32p_ok(direction, x, xcs, y, ycs)) {
            // the persistent kernel needs one CTA per SM co-resident (cooperative launch); where the context cannot give
            // that, the per-pass kernels do the same work
            rc = run_v32p(d, direction, x, xcs, y, ycs, M, ws, ws_bytes, st);
            if (rc != FMB_ERR_FALLBACK) return rc;
        }
        if (v32_ok(sizeof(C))) return run_v32(d, direction, x, xcs, y, ycs, M, ws, st);
        return run_fast<C>(d, direction, x, xcs, y, ycs, M, ws, st);
    }
    if (ws_bytes < slab * L * (int64_t)sizeof(C) || ws == nullptr) {
        set_error("workspace too small: need %lld bytes", (long long)(slab * L * (int64_t)sizeof(C)));
        return FMB_ERR_WORKSPACE;
    }
    const int64_t R1 = shape.g[0].R, R2 = shape.g[1].R;
    const bool kron = kron_a > 0;
    for (int64_t c0 = 0; c0 < M; c0 += slab) {
        const int64_t nc = std::min(slab, M - c0);
        const char *xs = (const char *)x + (size_t)(c0 * xcs) * (in_real ? sizeof(typename real_of<C>::type) : sizeof(C));
        char *ys = (char *)y + (size_t)(c0 * ycs) * sizeof(C);
        const int64_t trs = rowmajor ? nc : 1, tcs = rowmajor ? 1 : L;     // intermediate follows the caller's layout class

        // ---- pass A: length R1 over the row index n = n1*R2 + n2 (stride R2), lines n2
        {
            PassParams<C> p = blank_params<C>();
            p.lines_total = nc * R2; p.I = R2; p.ncols = nc; p.line_c_fastest = rowmajor;
            p.in = xs; p.in_real = in_real; p.in_cs = xcs; p.in_rs = xrs; p.in_lf = R2; p.in_li = 1; p.in_n = rows_in;
            p.pre = pre_d; p.pre_conj = bwd;
            if (!two_ffts && bwd) p.in_conj = 1;
            p.out = ws; p.out_cs = tcs; p.out_rs = trs; p.out_n = L;
            if (kron) { p.out_lk = R2; p.out_li = 1; }              // natural order k1*b + i2, no twiddle
            else {
                p.out_lk = 1; p.out_li = R1;                          // tmp[n2][k1]
                p.twL = (const C *)d.twL.p; p.twH = (const C *)d.twH.p; p.tw_shift = d.tw_shift;
                p.tw_mask = (unsigned)(((int64_t)1 << d.tw_shift) - 1);
            }
            p.wR = (const C *)d.wR[0].p;
            const bool last_t = rowmajor || kron;
            if ((rc = launch_pass<C>(p, shape.g[0], pow2, true, last_t, last_t, st))) return rc;
        }
        if (!two_ffts) {
            // ---- pass B: length R2, lines k1; writes y
            PassParams<C> p = blank_params<C>();
            p.lines_total = nc * R1; p.I = R1; p.ncols = nc; p.line_c_fastest = rowmajor;
            p.in = ws; p.in_cs = tcs; p.in_rs = trs; p.in_n = L;
            p.out = ys; p.out_cs = ycs; p.out_rs = yrs; p.out_n = rows_out;
            bool ord;
            if (kron) { p.in_lf = 1; p.in_li = R2; p.out_lk = 1; p.out_li = R2; ord = rowmajor; }
            else { p.in_lf = R1; p.in_li = 1; p.out_lk = R1; p.out_li = 1; ord = true; }
            if (bwd) p.conj_a = 1;
            p.wR = (const C *)d.wR[1].p;
            if ((rc = launch_pass<C>(p, shape.g[1], pow2, ord, ord, ord, st))) return rc;
        } else {
            // ---- pass B': FFT over n2 -> multiply spectrum -> conj -> FFT over k2, in place on the intermediate
            {
                PassParams<C> p = blank_params<C>();
                p.two_ffts = 1;
                p.lines_total = nc * R1; p.I = R1; p.ncols = nc; p.line_c_fastest = rowmajor;
                p.in = ws; p.in_cs = tcs; p.in_rs = trs; p.in_lf = R1; p.in_li = 1; p.in_n = L;
                p.mid = mid_d; p.mid_conj = bwd; p.mid_lk = 1; p.mid_li = R2;
                p.out = ws; p.out_cs = tcs; p.out_rs = trs; p.out_lk = R1; p.out_li = 1; p.out_n = L;
                p.twL = (const C *)d.twL.p; p.twH = (const C *)d.twH.p; p.tw_shift = d.tw_shift;
                p.tw_mask = (unsigned)(((int64_t)1 << d.tw_shift) - 1);
                p.wR = (const C *)d.wR[1].p;
                if ((rc = launch_pass<C>(p, shape.g[1], pow2, true, true, true, st))) return rc;
            }
            // ---- pass C: length R1 over k1, lines m2; conj (finishes the inverse), post-multiply, truncate, write y
            {
                PassParams<C> p = blank_params<C>();
                p.lines_total = nc * R2; p.I = R2; p.ncols = nc; p.line_c_fastest = rowmajor;
                p.in = ws; p.in_cs = tcs; p.in_rs = trs; p.in_lf = 1; p.in_li = R1; p.in_n = L;
                p.out = ys; p.out_cs = ycs; p.out_rs = yrs; p.out_lk = R2; p.out_li = 1; p.out_n = rows_out;
                p.conj_a = 1;
                p.post = post_d; p.post_conj = bwd;
                p.wR = (const C *)d.wR[0].p;
                if ((rc = launch_pass<C>(p, shape.g[0], pow2, rowmajor, rowmajor, true, st))) return rc;
            }
        }
    }
    return FMB_OK;
}

int ConvEngine::run(int direction, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs, int64_t ycs, int64_t M,
                    int dt_in, int dt_out, void *ws, int64_t ws_bytes, cudaStream_t st) const {
    if (dt_out == FMB_COMPLEX64 && (dt_in == FMB_COMPLEX64 || dt_in == FMB_FLOAT32))
        return run_t<float2>(dev_f, direction, x, xrs, xcs, dt_in == FMB_FLOAT32, y, yrs, ycs, M, ws, ws_bytes, st);
    if (dt_out == FMB_COMPLEX128 && (dt_in == FMB_COMPLEX128 || dt_in == FMB_FLOAT64))
        return run_t<double2>(dev_d, direction, x, xrs, xcs, dt_in == FMB_FLOAT64, y, yrs, ycs, M, ws, ws_bytes, st);
    set_error("FFT engine: unsupported dtype pair in=%d out=%d (cast the input to the output precision first)", dt_in, dt_out);
    return FMB_ERR_TYPE;
}

int device_fft_c128(const std::vector<cd> &in, std::vector<cd> &out) {
    ConvEngine e;
    int rc = e.init((int64_t)in.size(), (int64_t)in.size(), (int64_t)in.size(), false);
    if (rc) return rc;
    DevArray dx, dy, ws;
    if ((rc = dx.upload(in.data(), in.size() * sizeof(cd)))) return rc;
    if ((rc = dy.alloc(in.size() * sizeof(cd)))) return rc;
    int64_t wsb = e.workspace_bytes(1, sizeof(double2));
    if ((rc = ws.alloc((size_t)wsb))) return rc;
    rc = e.run(FMB_FORWARD, dx.p, 1, (int64_t)in.size(), dy.p, 1, (int64_t)in.size(), 1, FMB_COMPLEX128, FMB_COMPLEX128, ws.p, wsb, 0);
    if (rc) return rc;
    out.resize(in.size());
#ifndef FMB_EMULATE
    FMB_CUDA_OK(cudaStreamSynchronize(0));
#endif
    FMB_CUDA_OK(FMB_D2H(out.data(), dy.p, in.size() * sizeof(cd)));
    return FMB_OK;
}

}  // namespace fmb
