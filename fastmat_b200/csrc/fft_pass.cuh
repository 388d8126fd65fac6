// Generic batched FFT "pass" kernel: one CTA transforms a tile of T lines of length R that it stages in
// shared memory (Stockham autosort, radices 2/3/4/5/7/8/11/13/16, twiddles from a unit-root table), with
// everything that surrounds the transform in the reference's operator graph folded into its first load and
// last store:
//   load : zero-padding mask (Partial scatter / Toeplitz / Bluestein padding), real->complex, conj (backward),
//          pre-multiply (Bluestein chirp)
//   mid  : optional  FFT -> multiply by spectrum -> conj -> FFT  (the whole Product(F.H, Diag, F) chain of
//          fastmat/Circulant.pyx:131-133 without leaving shared memory)
//   store: four-step twiddle W_N^{i k}, conj, post-multiply, conj, scale, truncation mask (Partial gather)
// Large transforms are two such passes over an L2-resident intermediate (see fft_plan.cu).
//
// The body is __host__ __device__ and parameterised on the barrier so that tests/emul can execute the very same
// index logic with std::thread + std::barrier on the CPU (there is no GPU in the build container).
#pragma once
#include "cx.cuh"

namespace fmb {

constexpr int FMB_MAX_STAGES = 14;
constexpr int FMB_EMAX = 16;          // complex values a thread holds per stage
constexpr int FMB_MAX_NT = 512;       // threads per CTA (register budget 128/thread)

template <typename C> struct PassParams {
    typedef typename real_of<C>::type S;
    // ---- geometry
    int R;                       // transform length of this pass
    int T;                       // lines per tile
    int nstages;
    int radix[FMB_MAX_STAGES];
    int two_ffts;                // FFT -> (mid multiply, conj) -> FFT
    int t_fastest;               // thread order per stage position, bit0: stage reading global memory, bit1: inner
                                 // stages, bit2: stage writing global memory.  bit set: consecutive threads walk the
                                 // line index (lines contiguous in memory); clear: they walk the transform index
    int sf, st, psh, pamt;       // shared-memory index of (f, t): f*sf + t*st + (f >> psh)*pamt
    long long lines_total;       // ncols * I
    long long I;                 // lines per column
    long long ncols;
    int line_c_fastest;          // 0: line = c*I + i     1: line = i*ncols + c
    // ---- input: logical row n = f*in_lf + i*in_li, element at in[c*in_cs + n*in_rs], zero if n >= in_n
    const void *in;
    int in_real;
    long long in_cs, in_rs, in_lf, in_li, in_n;
    int in_conj;
    const C *pre; int pre_conj;  // indexed by n
    // ---- mid (two_ffts): multiply X[k] of line i by mid[k*mid_lk + i*mid_li] (conj if mid_conj), then conj
    const C *mid; int mid_conj;
    long long mid_lk, mid_li;
    // ---- output: logical row m = k*out_lk + i*out_li, element at out[c*out_cs + m*out_rs], dropped if m >= out_n
    void *out;
    long long out_cs, out_rs, out_lk, out_li, out_n;
    const C *twL; const C *twH; int tw_shift; unsigned tw_mask;   // W^{i*k} = twL[(i*k) & mask] * twH[(i*k) >> shift]
    int conj_a;
    const C *post; int post_conj;                                  // indexed by m
    int conj_b;
    S scale;
    const C *wR;                 // wR[j] = exp(-2 pi i j / R)
};

FMB_HD int ilog2(unsigned x) {
#ifdef __CUDA_ARCH__
    return 31 - __clz(x);
#else
    return 31 - __builtin_clz(x);
#endif
}

template <typename C> FMB_HD C ldg_c(const C *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// line index -> (column c, inner index i); all line arithmetic is 32-bit (the host guarantees lines_total < 2^31)
template <typename C, bool POW2>
FMB_HD bool decode_line(const PassParams<C> &p, unsigned tile, int t, unsigned &c, unsigned &i) {
    const unsigned line = tile * (unsigned)p.T + (unsigned)t;
    if (line >= (unsigned)p.lines_total) { c = 0; i = 0; return false; }
    if (p.line_c_fastest) { i = line / (unsigned)p.ncols; c = line - i * (unsigned)p.ncols; }
    else if (POW2) { const int sh = ilog2((unsigned)p.I); c = line >> sh; i = line & ((unsigned)p.I - 1u); }
    else { c = line / (unsigned)p.I; i = line - c * (unsigned)p.I; }
    return true;
}

template <typename C, bool POW2>
FMB_HD void decode_bfly(const PassParams<C> &p, int ord_bit, int Q, int b, int &j, int &t) {
    if (p.t_fastest & ord_bit) {
        if (POW2) { const int sh = ilog2((unsigned)p.T); t = b & (p.T - 1); j = b >> sh; }
        else { j = b / p.T; t = b - j * p.T; }
    } else {
        if (POW2) { const int sh = ilog2((unsigned)Q); j = b & (Q - 1); t = b >> sh; }
        else { t = b / Q; j = b - t * Q; }
    }
}

// SRC: 0 global, 1 shared, 2 shared * mid (then conj).   DST: 0 shared, 1 global.
// Deliberately NOT inlined on the device: each (radix, source, destination) variant gets its own register
// allocation (62-78 registers, no spills) instead of one allocation over the union of all variants.
template <typename C, int P, int SRC, int DST, bool POW2, typename Sync>
__host__ __device__ __noinline__ void run_stage(const PassParams<C> &p, int Ns, long long tile_ll, int tid, int NT, C *smem_arg, Sync &sync) {
    typedef typename real_of<C>::type S;
#ifdef __CUDA_ARCH__
    extern __shared__ __align__(16) unsigned char fmb_smem_raw[];      // keeps the accesses in the shared window (LDS/STS)
    C *smem = reinterpret_cast<C *>(fmb_smem_raw);
    (void)smem_arg;
#else
    C *smem = smem_arg;
#endif
    constexpr int NB = FMB_EMAX / P;
    constexpr int ORD = (SRC == 0) ? 1 : ((DST == 1) ? 4 : 2);
    const unsigned tile = (unsigned)tile_ll;
    const int Q = p.R / P;                       // butterflies per line
    const int nbfly = Q * p.T;
    const int tws = p.R / (Ns * P);              // twiddle stride in the unit-root table
    C v[NB][P];

#pragma unroll
    for (int it = 0; it < NB; ++it) {
        const int b = tid + it * NT;
        if (b < nbfly) {
            int j, t;
            decode_bfly<C, POW2>(p, ORD, Q, b, j, t);
            // ---------------- read
            if constexpr (SRC == 0) {
                unsigned c, i;
                const bool lv = decode_line<C, POW2>(p, tile, t, c, i);
                const long long base = (long long)c * p.in_cs + (long long)i * p.in_li * p.in_rs;
                const long long step = p.in_lf * p.in_rs;
                const long long nbase = (long long)i * p.in_li;
#pragma unroll
                for (int r = 0; r < P; ++r) {
                    const int f = j + r * Q;
                    const long long n = (long long)f * p.in_lf + nbase;
                    C val = mk<C>(0, 0);
                    if (lv && n < p.in_n) {
                        const long long a = base + (long long)f * step;
                        if (p.in_real) val = mk<C>(((const S *)p.in)[a], 0);
                        else val = ((const C *)p.in)[a];
                        if (p.in_conj) val = cconj(val);
                        if (p.pre) { C w = ldg_c(p.pre + n); val = p.pre_conj ? cmulc(val, w) : cmul(val, w); }
                    }
                    v[it][r] = val;
                }
            } else {
                long long mbase = 0;
                if constexpr (SRC == 2) {
                    unsigned c, i;
                    decode_line<C, POW2>(p, tile, t, c, i);
                    mbase = (long long)i * p.mid_li;
                }
#pragma unroll
                for (int r = 0; r < P; ++r) {
                    const int f = j + r * Q;
                    C val = smem[f * p.sf + t * p.st + (f >> p.psh) * p.pamt];
                    if constexpr (SRC == 2) {
                        C w = ldg_c(p.mid + ((long long)f * p.mid_lk + mbase));
                        val = p.mid_conj ? cmulc(val, w) : cmul(val, w);
                        val = cconj(val);
                    }
                    v[it][r] = val;
                }
            }
            // ---------------- inter-stage twiddles W_{Ns*P}^{r*(j mod Ns)}
            if (Ns > 1) {
                const int k = POW2 ? (j & (Ns - 1)) : (j % Ns);
                const int base = k * tws;
#pragma unroll
                for (int r = 1; r < P; ++r) v[it][r] = cmul(v[it][r], ldg_c(p.wR + base * r));
            }
        }
    }
    if constexpr (SRC != 0 && DST == 0) sync();            // in-place: every read of this stage precedes every write

    C wp[(P > 2 ? (P - 1) / 2 : 1) + 1];
    if constexpr (P == 5 || P == 7 || P == 11 || P == 13) {
#pragma unroll
        for (int r = 1; r <= (P - 1) / 2; ++r) wp[r] = ldg_c(p.wR + (long long)r * (p.R / P));
    }
#pragma unroll
    for (int it = 0; it < NB; ++it) {
        const int b = tid + it * NT;
        if (b < nbfly) {
            if constexpr (P == 2) dft2(v[it][0], v[it][1]);
            else if constexpr (P == 3) dft3(v[it][0], v[it][1], v[it][2]);
            else if constexpr (P == 4) dft4(v[it][0], v[it][1], v[it][2], v[it][3]);
            else if constexpr (P == 8) dft8(v[it]);
            else if constexpr (P == 16) dft16(v[it]);
            else dft_odd<C, P>(v[it], wp);
            int j, t;
            decode_bfly<C, POW2>(p, ORD, Q, b, j, t);
            int j0;
            if (POW2) j0 = ((j & ~(Ns - 1)) * P) + (j & (Ns - 1));
            else j0 = (j / Ns) * Ns * P + (j % Ns);
            if constexpr (DST == 0) {
#pragma unroll
                for (int q = 0; q < P; ++q) {
                    const int kk = j0 + q * Ns;
                    smem[kk * p.sf + t * p.st + (kk >> p.psh) * p.pamt] = v[it][outpos<P>(q)];
                }
            } else {
                unsigned c, i;
                if (decode_line<C, POW2>(p, tile, t, c, i)) {
                    const long long base = (long long)c * p.out_cs + (long long)i * p.out_li * p.out_rs;
                    const long long step = p.out_lk * p.out_rs;
                    const long long mbase = (long long)i * p.out_li;
#pragma unroll
                    for (int q = 0; q < P; ++q) {
                        const int kk = j0 + q * Ns;
                        const long long m = (long long)kk * p.out_lk + mbase;
                        if (m < p.out_n) {
                            C val = v[it][outpos<P>(q)];
                            if (p.twL) {
                                const unsigned long long e = (unsigned long long)i * (unsigned long long)kk;
                                C w = cmul(ldg_c(p.twL + (unsigned)(e & p.tw_mask)), ldg_c(p.twH + (unsigned)(e >> p.tw_shift)));
                                val = cmul(val, w);
                            }
                            if (p.conj_a) val = cconj(val);
                            if (p.post) { C w = ldg_c(p.post + m); val = p.post_conj ? cmulc(val, w) : cmul(val, w); }
                            if (p.conj_b) val = cconj(val);
                            if (p.scale != (S)1) val = cscale(val, p.scale);
                            ((C *)p.out)[base + (long long)kk * step] = val;
                        }
                    }
                }
            }
        }
    }
    if constexpr (DST == 0) sync();
}

template <typename C, int SRC, int DST, bool POW2, typename Sync>
FMB_HD void dispatch_stage(int P, const PassParams<C> &p, int Ns, long long tile, int tid, int NT, C *smem, Sync &sync) {
    switch (P) {
        case 16: run_stage<C, 16, SRC, DST, POW2, Sync>(p, Ns, tile, tid, NT, smem, sync); break;
        case 8: run_stage<C, 8, SRC, DST, POW2, Sync>(p, Ns, tile, tid, NT, smem, sync); break;
        case 4: run_stage<C, 4, SRC, DST, POW2, Sync>(p, Ns, tile, tid, NT, smem, sync); break;
        case 2: run_stage<C, 2, SRC, DST, POW2, Sync>(p, Ns, tile, tid, NT, smem, sync); break;
        default:
            if constexpr (!POW2) {
                switch (P) {
                    case 3: run_stage<C, 3, SRC, DST, false, Sync>(p, Ns, tile, tid, NT, smem, sync); break;
                    case 5: run_stage<C, 5, SRC, DST, false, Sync>(p, Ns, tile, tid, NT, smem, sync); break;
                    case 7: run_stage<C, 7, SRC, DST, false, Sync>(p, Ns, tile, tid, NT, smem, sync); break;
                    case 11: run_stage<C, 11, SRC, DST, false, Sync>(p, Ns, tile, tid, NT, smem, sync); break;
                    case 13: run_stage<C, 13, SRC, DST, false, Sync>(p, Ns, tile, tid, NT, smem, sync); break;
                    default: break;
                }
            }
            break;
    }
}

// The whole pass for one tile.  Every thread of the CTA calls this with the same arguments except `tid`.
template <typename C, bool POW2, typename Sync>
FMB_HD void pass_body(const PassParams<C> &p, long long tile, int tid, int NT, C *smem, Sync &sync) {
    const int S_ = p.nstages;
    int Ns = 1;
    if (!p.two_ffts) {
        for (int s = 0; s < S_; ++s) {
            const int P = p.radix[s];
            const bool first = (s == 0), last = (s == S_ - 1);
            if (first && last) dispatch_stage<C, 0, 1, POW2, Sync>(P, p, Ns, tile, tid, NT, smem, sync);
            else if (first) dispatch_stage<C, 0, 0, POW2, Sync>(P, p, Ns, tile, tid, NT, smem, sync);
            else if (last) dispatch_stage<C, 1, 1, POW2, Sync>(P, p, Ns, tile, tid, NT, smem, sync);
            else dispatch_stage<C, 1, 0, POW2, Sync>(P, p, Ns, tile, tid, NT, smem, sync);
            Ns *= P;
        }
    } else {
        for (int s = 0; s < S_; ++s) {
            const int P = p.radix[s];
            if (s == 0) dispatch_stage<C, 0, 0, POW2, Sync>(P, p, Ns, tile, tid, NT, smem, sync);
            else dispatch_stage<C, 1, 0, POW2, Sync>(P, p, Ns, tile, tid, NT, smem, sync);
            Ns *= P;
        }
        Ns = 1;
        for (int s = 0; s < S_; ++s) {
            const int P = p.radix[s];
            const bool first = (s == 0), last = (s == S_ - 1);
            if (first && last) dispatch_stage<C, 2, 1, POW2, Sync>(P, p, Ns, tile, tid, NT, smem, sync);
            else if (first) dispatch_stage<C, 2, 0, POW2, Sync>(P, p, Ns, tile, tid, NT, smem, sync);
            else if (last) dispatch_stage<C, 1, 1, POW2, Sync>(P, p, Ns, tile, tid, NT, smem, sync);
            else dispatch_stage<C, 1, 0, POW2, Sync>(P, p, Ns, tile, tid, NT, smem, sync);
            Ns *= P;
        }
    }
}

}  // namespace fmb
