// ConvEngine: the one device pipeline behind Fourier, Circulant, Toeplitz and Kron(Fourier, Fourier):
//
//      y = post . [ IFFT_L ( mid . ] FFT_L ( pad_L ( pre . x ) ) [ ) ]  restricted to n_out rows
//
// with every bracketed / dotted piece optional.  A plain Fourier is the un-bracketed form; Circulant and Toeplitz
// are the bracketed form with mid = fft(generator)/L (fastmat/Circulant.pyx:131, fastmat/Toeplitz.pyx:253-279)
// and the zero-padding / truncation of their Partial base class (fastmat/Partial.pyx:268-294) folded into the
// loads and stores; Bluestein (fastmat/Fourier.pyx:215-231) adds the chirp as pre and post.  backward() is the
// adjoint of the same pipeline (pre/post swapped and conjugated, mid conjugated).
#pragma once
#include <mutex>

#include "common.h"

namespace fmb {

struct PassGeom {
    int R = 0;
    std::vector<int> radix;
    int min_pnb = 16;          // min over stages of P * floor(16 / P): butterfly slots one thread covers
};

struct FftShape {
    int64_t L = 0;
    int npass = 0;             // 1 or 2
    PassGeom g[2];             // g[0]: first pass (length R1, stride R2), g[1]: second pass (length R2)
    bool pow2 = false;
};

// true if L can be transformed directly (all prime factors <= 13 and it splits into <= 2 shared-memory passes)
bool plan_shape(int64_t L, FftShape &shape, bool prefer_two = false);   // prefer_two: 8192 as 64 x 128 instead of one pass
int64_t next_pow2(int64_t v);

struct ConvEngine {
    int64_t L = 0;
    int64_t n_in = 0, n_out = 0;       // forward: rows read / rows written (backward swaps them)
    bool two_ffts = false;
    int64_t kron_a = 0, kron_b = 0;    // if > 0: 2-D transform Kron(Fourier(a), Fourier(b)) (no four-step twiddle)
    std::vector<cd> pre, post, mid;    // host masters (complex128); empty = absent
    FftShape shape;

    struct Dev {
        DevArray wR[2], twF[2], twV[2], twS32[2], twL, twH, twS[2], pre, post, mid;
        int tw_shift = 0;
        bool ready = false;
    };
    mutable Dev dev_f, dev_d;          // complex64 / complex128 device constants, built on first use
    mutable std::mutex mu;

    int init(int64_t L_, int64_t n_in_, int64_t n_out_, bool two_ffts_);
    int init_kron(int64_t a, int64_t b);
    int slab_cols(int64_t M, size_t csize) const;
    void slab_plan(int64_t M, size_t csize, bool fast, int &cols, int &ns) const;
    int64_t workspace_bytes(int64_t M, size_t csize) const;
    int64_t workspace_bytes_cm(int64_t M, size_t csize) const;      // column-major operands only
    int passes() const { return shape.npass == 1 ? 1 : (two_ffts ? 3 : 2); }

    // dt_in: FMB_FLOAT32/64 (real input) or FMB_COMPLEX64/128; dt_out: FMB_COMPLEX64/128 (same precision as dt_in)
    int run(int direction, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs, int64_t ycs, int64_t M, int dt_in,
            int dt_out, void *ws, int64_t ws_bytes, cudaStream_t st) const;

   private:
    template <typename C> int ensure_dev(Dev &d) const;
    template <typename C> bool fast_ok(int64_t xrs, int64_t yrs, bool in_real) const;
    template <typename C>
    int run_single_fast(Dev &d, int direction, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs, int64_t ycs, int64_t M,
                        int64_t &done, cudaStream_t st) const;
    template <typename C>
    int run_fast(Dev &d, int direction, const void *x, int64_t xcs, void *y, int64_t ycs, int64_t M, void *ws, cudaStream_t st) const;
    bool v32_ok(size_t csize) const;
    int run_v32(Dev &d, int direction, const void *x, int64_t xcs, void *y, int64_t ycs, int64_t M, void *ws, cudaStream_t st) const;
    bool v32p_ok(int direction, const void *x, int64_t xcs, const void *y, int64_t ycs) const;
    int run_v32p(Dev &d, int direction, const void *x, int64_t xcs, void *y, int64_t ycs, int64_t M, void *ws, int64_t ws_bytes,
                 cudaStream_t st) const;
    int64_t v32p_workspace_bytes(int64_t M) const;
    template <typename C>
    int run_t(Dev &d, int direction, const void *x, int64_t xrs, int64_t xcs, bool in_real, void *y, int64_t yrs, int64_t ycs,
              int64_t M, void *ws, int64_t ws_bytes, cudaStream_t st) const;
};

// forward complex128 DFT of a host vector on the device (used at plan creation for spectra); synchronous
int device_fft_c128(const std::vector<cd> &in, std::vector<cd> &out);

}  // namespace fmb
