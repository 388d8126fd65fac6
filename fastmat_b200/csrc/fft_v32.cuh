// "V32" passes: the length-1024 pass of a 2^20-point transform with THIRTY-TWO values per thread.
//
// 1024 = 32 * 32: a thread owns the 32 positions jb + 32 m of one line, so a pass is two radix-32 stages with ONE
// exchange through shared memory (the 16-values-per-thread path of fft_fast.cuh needs 16 * 16 * 4: two exchanges), and
// because the last stage of one transform leaves thread jb with exactly the positions jb + 32 q that the first stage of
// the next transform reads, the middle pass of a convolution (FFT -> * spectrum -> conj -> FFT) needs two exchanges
// instead of five.  A line is owned by one warp in the row-fastest order: the middle pass, whose lines are contiguous
// in the (transposed) intermediate, contains no CTA-wide barrier at all - only __syncwarp().
//
// Shared-memory traffic per point drops from 25 to 15 accesses over the three passes of a Circulant apply; that pipe
// (128 B/clk/SM) and the issue slots are the two co-limiters of these kernels.
//
// Same argument block and option bits as the fast path (FastArgs / FO_*); 256 threads, 8 lines per tile, <= 128
// registers, 2 CTAs per SM.  complex64 only (complex128 stays on the 16-value path: 32 double2 values do not fit).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fft_fast.cuh"

namespace fmb {

constexpr int V32_LOGT = 3, V32_T = 1 << V32_LOGT, V32_NT = 256;
constexpr int V32_RS = 1058;                   // line stride: 1024 + one pad per 32, == 2 (mod 16) (see fft_fast.cuh)
constexpr size_t V32_SMEM = (size_t)V32_T * V32_RS * sizeof(float2);

// z * exp(-2 pi i K / 32)
template <int K, typename C> __device__ __forceinline__ C mul_w32(C z) {
    typedef typename real_of<C>::type S;
    if constexpr (K == 0) return z;
    else if constexpr (K == 8) return cmul_mi(z);
    else if constexpr (K == 4) { const S h = (S)0.70710678118654752440084436210485; return mk<C>((z.x + z.y) * h, (z.y - z.x) * h); }
    else if constexpr (K == 12) { const S h = (S)0.70710678118654752440084436210485; return mk<C>((z.y - z.x) * h, -(z.x + z.y) * h); }
    else {
        constexpr double ang = 6.283185307179586476925286766559 * K / 32.0;
        // constexpr cos / sin of the sixteen angles (no constexpr libm in device code): table
        constexpr double ct[16] = {1.0, 0.98078528040323044912618223613424, 0.92387953251128675612818318939679,
                                   0.83146961230254523707878837761791, 0.70710678118654752440084436210485,
                                   0.55557023301960222474283081394853, 0.38268343236508977172845998403040,
                                   0.19509032201612826784828486847702, 0.0, -0.19509032201612826784828486847702,
                                   -0.38268343236508977172845998403040, -0.55557023301960222474283081394853,
                                   -0.70710678118654752440084436210485, -0.83146961230254523707878837761791,
                                   -0.92387953251128675612818318939679, -0.98078528040323044912618223613424};
        constexpr double st[16] = {0.0, 0.19509032201612826784828486847702, 0.38268343236508977172845998403040,
                                   0.55557023301960222474283081394853, 0.70710678118654752440084436210485,
                                   0.83146961230254523707878837761791, 0.92387953251128675612818318939679,
                                   0.98078528040323044912618223613424, 1.0, 0.98078528040323044912618223613424,
                                   0.92387953251128675612818318939679, 0.83146961230254523707878837761791,
                                   0.70710678118654752440084436210485, 0.55557023301960222474283081394853,
                                   0.38268343236508977172845998403040, 0.19509032201612826784828486847702};
        (void)ang;
        const S c = (S)ct[K], s = (S)st[K];
        return mk<C>(z.x * c + z.y * s, z.y * c - z.x * s);
    }
}

// cos / sin of 2 pi K / 32 as compile-time constants (K < 16)
template <int K> struct W32 {
    static constexpr double c[16] = {1.0, 0.98078528040323044912618223613424, 0.92387953251128675612818318939679,
                                     0.83146961230254523707878837761791, 0.70710678118654752440084436210485,
                                     0.55557023301960222474283081394853, 0.38268343236508977172845998403040,
                                     0.19509032201612826784828486847702, 0.0, -0.19509032201612826784828486847702,
                                     -0.38268343236508977172845998403040, -0.55557023301960222474283081394853,
                                     -0.70710678118654752440084436210485, -0.83146961230254523707878837761791,
                                     -0.92387953251128675612818318939679, -0.98078528040323044912618223613424};
    static constexpr double s[16] = {0.0, 0.19509032201612826784828486847702, 0.38268343236508977172845998403040,
                                     0.55557023301960222474283081394853, 0.70710678118654752440084436210485,
                                     0.83146961230254523707878837761791, 0.92387953251128675612818318939679,
                                     0.98078528040323044912618223613424, 1.0, 0.98078528040323044912618223613424,
                                     0.92387953251128675612818318939679, 0.83146961230254523707878837761791,
                                     0.70710678118654752440084436210485, 0.55557023301960222474283081394853,
                                     0.38268343236508977172845998403040, 0.19509032201612826784828486847702};
    static constexpr double cosv = c[K], sinv = s[K];
};

// X[K] = E[K] + W32^K O[K],  X[K + 16] = E[K] - W32^K O[K]: the twiddle is folded into the butterfly (bfly_w, 6 fused
// operations instead of 4 + 4); K = 0 and K = 8 (W = -i) need no multiplication at all
template <typename C, int K> __device__ __forceinline__ void dft32_combine(C (&v)[32], const C (&e)[16], const C (&o)[16]) {
    typedef typename real_of<C>::type S;
    if constexpr (K < 16) {
        constexpr int p = outpos<16>(K);
#ifdef FMB_PLAIN_BUTTERFLIES
        const C t = mul_w32<K>(o[p]);
        v[K] = cadd(e[p], t);
        v[K + 16] = csub(e[p], t);
#else
        if constexpr (K == 0) { v[0] = cadd(e[p], o[p]); v[16] = csub(e[p], o[p]); }
        else if constexpr (K == 8) {
            v[8] = mk<C>(e[p].x + o[p].y, e[p].y - o[p].x);
            v[24] = mk<C>(e[p].x - o[p].y, e[p].y + o[p].x);
        } else bfly_w(e[p], o[p], (S)W32<K>::cosv, (S)-W32<K>::sinv, v[K], v[K + 16]);
#endif
        dft32_combine<C, K + 1>(v, e, o);
    }
}

// forward 32-point DFT in registers, natural order in and out: two radix-16 butterflies on the even / odd inputs and
// a radix-2 combine  X[k] = E[k] + W32^k O[k],  X[k+16] = E[k] - W32^k O[k]
template <typename C> __device__ __forceinline__ void dft32(C (&v)[32]) {
    C e[16], o[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) { e[r] = v[2 * r]; o[r] = v[2 * r + 1]; }
    dft16(e);
    dft16(o);
    dft32_combine<C, 0>(v, e, o);
}

// dft32 of 32 inputs whose upper half (m >= 16) is zero (zero padding of a Toeplitz / linear convolution to twice the
// length): decimation in frequency skips the first radix-2 level - even outputs are the 16-point DFT of the inputs, odd
// outputs the 16-point DFT of the inputs times W32^m.  Natural order in and out, like dft32.
template <typename C, int M> __device__ __forceinline__ void dft32_uz_prep(const C (&v)[32], C (&e)[16], C (&o)[16]) {
    if constexpr (M < 16) {
        e[M] = v[M];
        o[M] = mul_w32<M>(v[M]);
        dft32_uz_prep<C, M + 1>(v, e, o);
    }
}
template <typename C> __device__ __forceinline__ void dft32_upper_zero(C (&v)[32]) {
    C e[16], o[16];
    dft32_uz_prep<C, 0>(v, e, o);
    dft16(e);
    dft16(o);
#pragma unroll
    for (int r = 0; r < 16; ++r) { v[2 * r] = e[outpos<16>(r)]; v[2 * r + 1] = o[outpos<16>(r)]; }
}
// dft32 whose outputs q >= 16 are not needed (the rows a Toeplitz / linear convolution discards): the radix-2 combine
// produces only E[k] + W32^k O[k].
template <typename C, int K> __device__ __forceinline__ void dft32_combine_lower(C (&v)[32], const C (&e)[16], const C (&o)[16]) {
    if constexpr (K < 16) {
        constexpr int p = outpos<16>(K);
        typedef typename real_of<C>::type S;
        if constexpr (K == 0) v[0] = cadd(e[p], o[p]);
        else if constexpr (K == 8) v[8] = mk<C>(e[p].x + o[p].y, e[p].y - o[p].x);
        else v[K] = add_w(e[p], o[p], (S)W32<K>::cosv, (S)-W32<K>::sinv);
        dft32_combine_lower<C, K + 1>(v, e, o);
    }
}
template <typename C> __device__ __forceinline__ void dft32_lower_only(C (&v)[32]) {
    C e[16], o[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) { e[r] = v[2 * r]; o[r] = v[2 * r + 1]; }
    dft16(e);
    dft16(o);
    dft32_combine_lower<C, 0>(v, e, o);
}

// stage-twiddle pair load that is neither merged with the identical load of the pass's other transform nor hoisted as a
// block (the compiler would otherwise keep all sixteen pairs - 64 registers - alive across the whole middle pass)
__device__ __forceinline__ CPair<float2> ldg_pair_ordered(const CPair<float2> *p) {
    CPair<float2> r;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.b.x), "=f"(r.b.y) : "l"(p));
    return r;
}

// Second radix-32 stage of a 1024-point transform: dft32 of v[m] * W_1024^{jb m}.  The 31 stage twiddles are not multiplied
// in beforehand (124 operations) but ride on the first butterfly level of the two 16-point sub-transforms (dft4_tw4:
// 4 operations less per radix-4 butterfly, 32 per call): the four pairs {W^{jb 2p}, W^{jb (2p+1)}}, p = r0, 4 + r0, 8 + r0,
// 12 + r0, that the r0-th butterflies of the even and the odd sub-transform need are loaded right before them.
// `tp` = table + jb (pairs 32 apart).  LOWER: outputs q >= 16 are not needed.
template <typename C, bool LOWER = false> __device__ __forceinline__ void dft32_stage_tw(C (&v)[32], const CPair<C> *tp) {
    C e[16], o[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) { e[r] = v[2 * r]; o[r] = v[2 * r + 1]; }
#pragma unroll
    for (int r0 = 0; r0 < 4; ++r0) {
        const CPair<C> w0 = ldg_pair_ordered(tp + r0 * 32), w1 = ldg_pair_ordered(tp + (4 + r0) * 32),
                       w2 = ldg_pair_ordered(tp + (8 + r0) * 32), w3 = ldg_pair_ordered(tp + (12 + r0) * 32);
        if (r0 == 0) dft4_tw(e[0], e[4], e[8], e[12], w1.a.x, w1.a.y, w2.a.x, w2.a.y, w3.a.x, w3.a.y);     // m = 0: twiddle 1
        else dft4_tw4(e[r0], e[4 + r0], e[8 + r0], e[12 + r0], w0.a, w1.a, w2.a, w3.a);
        dft4_tw4(o[r0], o[4 + r0], o[8 + r0], o[12 + r0], w0.b, w1.b, w2.b, w3.b);
    }
    dft16_level2(e);
    dft16_level2(o);
    if constexpr (LOWER) dft32_combine_lower<C, 0>(v, e, o);
    else dft32_combine<C, 0>(v, e, o);
}

// dft32 whose odd inputs are prepared by `prep_odd()` only after the even half has been transformed: the middle pass
// uses it to overlap the latency of the second half of its spectrum loads with sixteen-point butterflies
template <typename C, typename PrepOdd> __device__ __forceinline__ void dft32_late_odd(C (&v)[32], PrepOdd prep_odd) {
    C e[16], o[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) e[r] = v[2 * r];
    dft16(e);
    prep_odd();
#pragma unroll
    for (int r = 0; r < 16; ++r) o[r] = v[2 * r + 1];
    dft16(o);
    dft32_combine<C, 0>(v, e, o);
}

template <bool ORDER_T, int LOGT = 3> __device__ __forceinline__ void v32_pos(int tid, int &jb, int &t) {
    if (ORDER_T) { t = tid & ((1 << LOGT) - 1); jb = tid >> LOGT; }
    else { jb = tid & 31; t = tid >> 5; }
}

// warp-level barrier if the line stays inside one warp on both sides (row-fastest before and after), else CTA-wide
template <bool WARP> __device__ __forceinline__ void v32_sync() {
    if constexpr (WARP) __syncwarp();
    else __syncthreads();
}

// Four-step twiddle W_L^{i (jb + 32 q)}, q = 0..31, of line i as four interleaved chains: c[b] = W^{i jb} s^b (b < 4) with
// s = W_L^{32 i}, each advanced by s4 = s^4 after use (value q is c[q & 3] at the time it is used).
struct V32Chain { float2 c[4], s4; };
__device__ __forceinline__ V32Chain v32_chain_init(const FastArgs<float2> &a, unsigned i, int jb) {
    typedef float2 C;
    V32Chain h;
    const unsigned e = i * (unsigned)jb;
    h.c[0] = cmul(__ldg(a.twL + (e & a.tw_mask)), __ldg(a.twH + (e >> a.tw_shift)));
    const C s1 = __ldg(a.twS + i);
    const C s2 = cmul(s1, s1);
    h.c[1] = cmul(h.c[0], s1);
    h.c[2] = cmul(h.c[0], s2);
    h.c[3] = cmul(h.c[1], s2);
    h.s4 = cmul(s2, s2);
    return h;
}

#ifdef V32_TIMING
std::vector<void (*)()> &debug_dumpers();        // capi.cu; run by fmb_debug_dump()
// Phase timing (experiment builds only, -DV32_TIMING): lane 0 of every warp of every 16th CTA records the SM clock at the
// phase boundaries; the clock read is predicated on a value the phase produced, so it cannot be scheduled before it.
constexpr int V32_TREC = 1 << 16, V32_TFIELDS = 8;
__device__ __forceinline__ unsigned long long v32_clock_after(float dep) {
    unsigned long long t;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.neu.f32 p, %1, %1;\n\t@!p mov.u64 %0, %%clock64;\n\t@p mov.u64 %0, 0;\n\t}" : "=l"(t) : "f"(dep) : "memory");
    return t;
}
#define V32_T_DECL unsigned long long tph[V32_TFIELDS] = {}; int tn = 0; const bool trec = tbuf != nullptr && (blockIdx.x & 15) == 0 && (threadIdx.x & 31) == 0;
#define V32_T_MARK(dep) do { if (trec) tph[tn] = v32_clock_after(dep); ++tn; } while (0)
#else
#define V32_T_DECL
#define V32_T_MARK(dep) do { } while (0)
#endif

// Geometry is fixed (L = 1024 x 1024, everything but the column stride): line-fastest sides address element (f, i) of a
// column at f * 1024 + i, row-fastest sides at i * 1024 + f, so that a thread's 32 loads / stores are ONE base register
// plus immediates (the run-time strides of the general FastArgs cost ~6 integer instructions per access: a fifth of the
// instructions of the strided passes).  launch_v32_variant() checks that the arguments describe exactly this geometry.
template <unsigned OPT, int LOGT, int MINB>
__global__ void __launch_bounds__(32 << LOGT, MINB) v32_pass_kernel(const __grid_constant__ FastArgs<float2> a
#ifdef V32_TIMING
                                                                      , unsigned long long *tbuf, unsigned *tcount
#endif
) {
    typedef float2 C;
    V32_T_DECL
    extern __shared__ __align__(16) unsigned char fmb_v32_smem[];
    C *const smem = reinterpret_cast<C *>(fmb_v32_smem);
    constexpr bool LOAD_T = (OPT & FO_LOAD_T) != 0, STORE_T = (OPT & FO_STORE_T) != 0, TWO = (OPT & FO_TWO_FFTS) != 0;
    static_assert(LOGT == V32_LOGT || (!LOAD_T && !STORE_T), "strided sides need the 8-line tile (64-byte segments)");
    const int tid = threadIdx.x;
    // (one tile per CTA: two or four tiles per CTA, to halve the CTA launches and keep the L1 warm, measured 25 - 55 % slower)
    const unsigned line0 = blockIdx.x << LOGT;
    const unsigned col = line0 >> 10;                      // the tile width divides 1024: a tile never straddles two columns
    const unsigned i0 = line0 & 1023u;
    C v[32];
    int jb, t;

    V32_T_MARK(0.f);
    if constexpr (LOAD_T && STORE_T && (OPT & FO_TWIDDLE)) {
        // first pass: ask the L2 for the 64 KB of a LATER slab's input that correspond to this tile's share (the 128 tiles
        // of a column cover it in contiguous chunks), so that slab's loads find their data on chip (FMB_V32_PF, experiments)
        if (a.pf != nullptr && tid == 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.pf + (long long)col * a.in_cs + (long long)(i0 >> 3) * 8192), "r"(65536u) : "memory");
    }
    // Stage twiddles are read through L1 where they are needed.  Measured alternatives (round 2, profiles/r2_experiments.txt):
    // copying the 8 KB table into shared memory with cp.async at the start of the CTA shortens the second stage (the L1 is
    // cold after every launch boundary) but the copy competes with the tile's own loads: +4 % overall; requesting the
    // store side's chain start values up front costs registers during the load phase: +40 %.
    const CPair<C> *const tab = reinterpret_cast<const CPair<C> *>(a.tw);
    // ------------------------------------------------------------------ global -> registers, first radix-32 stage
    v32_pos<LOAD_T, LOGT>(tid, jb, t);
    {
        const unsigned i = i0 + t;
        const C *src = a.in + (long long)col * a.in_cs +
                       (LOAD_T ? (int)i + jb * 1024 : ((OPT & FO_DYN_LINES) ? (int)i * a.in_is + jb : (int)i * 1024 + jb));
        constexpr int fstep = 32 * 1024;
#ifdef FMB_PLAIN_BUTTERFLIES
        V32Chain h;
        if (OPT & FO_IN_TWIDDLE) h = v32_chain_init(a, i, jb);
#endif
        // zero padding: logical row f*in_lf + i*in_li < in_n  <=>  32 m < (number of valid f) - jb, one compare per value
        int flim = 0;
        if (OPT & FO_IN_MASK) {
            const int room = a.in_n - (int)i * a.in_li;
            flim = (room > 0 ? (room + a.in_lf - 1) / a.in_lf : 0) - jb;
        }
#pragma unroll
        for (int m = 0; m < 32; ++m) {
            const int f = jb + 32 * m;
            if ((OPT & FO_IN_HALF) && m >= 16) { v[m] = mk<C>(0, 0); continue; }
            bool ok = true;
            if (OPT & FO_IN_MASK) ok = 32 * m < flim;
            C val = mk<C>(0, 0);
            // the address is formed outside the predicate so that a masked load is one predicated LDG, not a branch
            const C *pl = LOAD_T ? src + m * fstep : src + 32 * m;      // row-fastest: the transform index is contiguous
            if (ok) {
#ifdef V32_DEBUG_NOLOAD                                                 /* timing experiment only: how much do the input loads cost? */
                val = mk<C>((float)(jb + m) * 1e-3f, (float)(t + (int)(reinterpret_cast<size_t>(pl) & 1)));
#elif !defined(V32_NO_STREAM_LOADS)
                val = ld_stream(pl);
#else
                val = *pl;
#endif
                if (OPT & FO_IN_CONJ) val = cconj(val);
                if (OPT & FO_PRE) {
                    const C w = __ldg(a.pre + (f * a.in_lf + (int)i * a.in_li));
                    val = (OPT & FO_PRE_CONJ) ? cmulc(val, w) : cmul(val, w);
                }
            }
#ifdef FMB_PLAIN_BUTTERFLIES
            if (OPT & FO_IN_TWIDDLE) {
                // the four-step twiddle of the PREVIOUS pass's output, applied here: the chain arithmetic runs while the
                // loads are in flight, and the (FP32-bound) middle pass of a convolution is relieved of it
                val = cmul(val, h.c[m & 3]);
                if (m + 4 < 32) h.c[m & 3] = cmul(h.c[m & 3], h.s4);
            }
#endif
            v[m] = val;
        }
    }
#ifdef V32_TIMING
    V32_T_MARK(v[0].x + v[31].y + v[16].x + v[15].y);                   // (most) loads have arrived, input-side multiplies done
#endif
    if constexpr ((OPT & FO_IN_HALF) != 0) dft32_upper_zero(v);
#ifndef FMB_PLAIN_BUTTERFLIES
    else if constexpr ((OPT & FO_IN_TWIDDLE) != 0) {
        // The four-step twiddle W_L^{i (jb + 32 m)} of the PREVIOUS pass's output is applied here (the FP32-bound middle
        // pass of a convolution is relieved of it), folded into the first butterfly level: the r0-th radix-4 butterflies of
        // the even / odd sub-transform take inputs m = b + 8 j (b = 2 r0 / 2 r0 + 1, j < 4), i.e. the chain
        // c_b (s^8)^j with c_b = W^{i jb} s^b, s = W_L^{32 i}.  Same number of chain multiplications as applying them one
        // by one, 32 operations less for the application.
        const unsigned i = i0 + t;
        const unsigned ee = i * (unsigned)jb;
        C cb = cmul(__ldg(a.twL + (ee & a.tw_mask)), __ldg(a.twH + (ee >> a.tw_shift)));       // W^{i jb}
        const C s1 = __ldg(a.twS + i);
        const C s2 = cmul(s1, s1), s4 = cmul(s2, s2), s8 = cmul(s4, s4);
        C e[16], o[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) { e[r] = v[2 * r]; o[r] = v[2 * r + 1]; }
#pragma unroll
        for (int r0 = 0; r0 < 4; ++r0) {
            const C w0 = cb, w1 = cmul(w0, s8), w2 = cmul(w1, s8), w3 = cmul(w2, s8);
            dft4_tw4(e[r0], e[4 + r0], e[8 + r0], e[12 + r0], w0, w1, w2, w3);
            cb = cmul(cb, s1);
            const C x0 = cb, x1 = cmul(x0, s8), x2 = cmul(x1, s8), x3 = cmul(x2, s8);
            dft4_tw4(o[r0], o[4 + r0], o[8 + r0], o[12 + r0], x0, x1, x2, x3);
            if (r0 < 3) cb = cmul(cb, s1);
        }
        dft16_level2(e);
        dft16_level2(o);
        dft32_combine<C, 0>(v, e, o);
    }
#endif
    else dft32(v);
    V32_T_MARK(v[0].x + v[31].y);                                       // first butterfly done
    auto exchange_store = [&](int jb_, int t_) {
        C *sl = smem + t_ * V32_RS + jb_ * 33;                          // position k = 32 jb + q at k + (k >> 5)
#pragma unroll
        for (int q = 0; q < 32; ++q) sl[q] = v[q];
    };
    exchange_store(jb, t);

    // second radix-32 stage: v[m] <- position jb + 32 m, times W_1024^{jb m}, DFT; leaves X[jb + 32 q] in v[q]
    auto stage_b = [&](int jb_, int t_, const CPair<C> *tb) {
        const C *sl = smem + t_ * V32_RS + jb_;
#pragma unroll
        for (int m = 0; m < 32; ++m) v[m] = sl[33 * m];
        const CPair<C> *tp = tb + jb_;
#if defined(V32_DEBUG_NOTW) || defined(FMB_PLAIN_BUTTERFLIES)
#pragma unroll
        for (int p2 = 0; p2 < 16; ++p2) {
#ifdef V32_DEBUG_NOTW                                                   /* timing experiment only */
            CPair<C> w; w.a = mk<C>(0.5f + (float)p2, 0.25f); w.b = mk<C>(0.75f, (float)jb_);
#else
            const CPair<C> w = ldg_pair_ordered(tp + p2 * 32);
#endif
            if (p2 > 0) v[2 * p2] = cmul(v[2 * p2], w.a);
            v[2 * p2 + 1] = cmul(v[2 * p2 + 1], w.b);
        }
        if constexpr ((OPT & FO_OUT_HALF) != 0 && !TWO) dft32_lower_only(v);
        else dft32(v);
#else
        dft32_stage_tw<C, (OPT & FO_OUT_HALF) != 0 && !TWO>(v, tp);
#endif
    };

    // last stage output -> global: four-step twiddle W^{i k}, conj, mask, post-multiply
    auto final_store = [&](int jb_, int t_) {
        const unsigned i = i0 + t_;
        C *dst = a.out + (long long)col * a.out_cs +
                 (STORE_T ? (int)i + jb_ * 1024 : ((OPT & FO_DYN_LINES) ? (int)i * a.out_is + jb_ : (int)i * 1024 + jb_));
        constexpr int kstep = 32 * 1024;
        V32Chain hst;
        if (OPT & FO_TWIDDLE) hst = v32_chain_init(a, i, jb_);
        int klim = 0;                                                    // truncation: k*out_lk + i*out_li < out_n
        if (OPT & FO_OUT_MASK) {
            const int room = a.out_n - (int)i * a.out_li;
            klim = (room > 0 ? (room + a.out_lk - 1) / a.out_lk : 0) - jb_;
        }
#pragma unroll
        for (int q = 0; q < ((OPT & FO_OUT_HALF) ? 16 : 32); ++q) {
            C val = v[q];
            if (OPT & FO_TWIDDLE) {
                val = cmul(val, hst.c[q & 3]);
                if (q + 4 < 32) hst.c[q & 3] = cmul(hst.c[q & 3], hst.s4);
            }
            if (OPT & FO_OUT_CONJ) val = cconj(val);
            const int k = jb_ + 32 * q;
            bool ok = true;
            const int mrow = k * a.out_lk + (int)i * a.out_li;
            if (OPT & FO_OUT_MASK) ok = 32 * q < klim;
            if (OPT & FO_POST) {
                if (ok) {
                    const C pw = __ldg(a.post + mrow);
                    val = (OPT & FO_POST_CONJ) ? cmulc(val, pw) : cmul(val, pw);
                }
            }
            C *ps = STORE_T ? dst + q * kstep : dst + 32 * q;          // row-fastest: the output index is contiguous
#ifdef V32_DEBUG_NOSTORE                                                /* timing experiment only */
            if (ok && val.x == 1.2345e30f) *ps = val;
#else
            if (ok) *ps = val;
#endif
        }
    };

    constexpr bool WARP_ONLY_1 = TWO ? !LOAD_T : (!LOAD_T && !STORE_T);   // lines stay inside one warp: __syncwarp() suffices
    if constexpr (!TWO) {
        v32_sync<WARP_ONLY_1>();
        V32_T_MARK(0.f);                                                // exchange stores issued, barrier passed
        v32_pos<STORE_T, LOGT>(tid, jb, t);
        stage_b(jb, t, tab);
        V32_T_MARK(v[0].x + v[31].y);                                   // second stage done
        final_store(jb, t);
        V32_T_MARK(0.f);                                                // stores issued
    } else {
        // ---- middle pass of a convolution.  Inner stages are row-fastest: lane jb of warp t owns line t.
        v32_sync<WARP_ONLY_1>();
        v32_pos<false, LOGT>(tid, jb, t);
        // Spectrum values of this thread's outputs k = jb + 32 q (one L2 round trip each): those of the even q are
        // requested now - the registers of v are free between the exchange store and load - and arrive during the second
        // stage; those of the odd q are requested once the even ones are consumed and arrive during the even half of the
        // next transform's first butterfly.  (Fetched a few at a time they cost eight exposed round trips per tile.)
        const C *mp = a.mid + (long long)(i0 + t) * a.mid_is + jb;
        C mh[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) mh[r] = ld_nc_ordered(mp + 32 * (2 * r));
        stage_b(jb, t, tab);
        V32_T_MARK(v[0].x + v[31].y);                                   // first transform done
#ifdef FMB_PLAIN_BUTTERFLIES
#pragma unroll
        for (int r = 0; r < 16; ++r) v[2 * r] = cconj((OPT & FO_MID_CONJ) ? cmulc(v[2 * r], mh[r]) : cmul(v[2 * r], mh[r]));
#pragma unroll
        for (int r = 0; r < 16; ++r) mh[r] = ld_nc_ordered(mp + 32 * (2 * r + 1));
        // v[q] = position jb + 32 q: exactly the input of the next transform's first stage - no exchange
        __syncwarp();                                                    // every lane has read its stage inputs
        dft32_late_odd(v, [&]() {
#pragma unroll
            for (int r = 0; r < 16; ++r)
                v[2 * r + 1] = cconj((OPT & FO_MID_CONJ) ? cmulc(v[2 * r + 1], mh[r]) : cmul(v[2 * r + 1], mh[r]));
        });
#else
        {
            // conj(X S) = conj(X) conj(S): the spectrum rides as a twiddle on the first butterfly level of the inverse
            // transform's first stage (dft4_tw4) instead of 32 separate complex multiplications.  v[q] = position jb + 32 q
            // is exactly that stage's input - no exchange.  Even half first; the odd half's spectrum values are requested
            // once the even ones are consumed and arrive during the even half's second level.
            auto sw = [](C m) { return (OPT & FO_MID_CONJ) ? m : cconj(m); };
            C e[16], o[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) e[r] = cconj(v[2 * r]);
#pragma unroll
            for (int r0 = 0; r0 < 4; ++r0)
                dft4_tw4(e[r0], e[4 + r0], e[8 + r0], e[12 + r0], sw(mh[r0]), sw(mh[4 + r0]), sw(mh[8 + r0]), sw(mh[12 + r0]));
#pragma unroll
            for (int r = 0; r < 16; ++r) mh[r] = ld_nc_ordered(mp + 32 * (2 * r + 1));
            __syncwarp();                                                // every lane has read its stage inputs
            dft16_level2(e);
#pragma unroll
            for (int r = 0; r < 16; ++r) o[r] = cconj(v[2 * r + 1]);
#pragma unroll
            for (int r0 = 0; r0 < 4; ++r0)
                dft4_tw4(o[r0], o[4 + r0], o[8 + r0], o[12 + r0], sw(mh[r0]), sw(mh[4 + r0]), sw(mh[8 + r0]), sw(mh[12 + r0]));
            dft16_level2(o);
            dft32_combine<C, 0>(v, e, o);
        }
#endif
        V32_T_MARK(v[0].x + v[31].y);                                   // spectrum product + first stage of the second transform
        exchange_store(jb, t);
        v32_sync<!STORE_T>();
        v32_pos<STORE_T, LOGT>(tid, jb, t);
        V32_T_MARK(0.f);
        // second copy of the table: the compiler merges loads from identical addresses with the first transform's and then
        // keeps all sixteen pairs (64 registers) alive across the pass - measured (round 2, by accident): 2.30 -> 3.25 ms
        stage_b(jb, t, tab + 512);
        V32_T_MARK(v[0].x + v[31].y);
        final_store(jb, t);
        V32_T_MARK(0.f);
    }
#ifdef V32_TIMING
    if (trec) {
        const unsigned r = atomicAdd(tcount, 1u);
        if (r < (unsigned)V32_TREC)
            for (int f = 0; f < V32_TFIELDS; ++f) tbuf[(size_t)r * V32_TFIELDS + f] = tph[f];
    }
#endif
}

// ---- pass variants used by the engine (fft_engine.cu: run_v32); the intermediate is [k1][n2] (n2 contiguous)
constexpr unsigned V32_A_F = FO_LOAD_T | FO_STORE_T | FO_TWIDDLE;                // first pass: strided in, strided out
constexpr unsigned V32_A_FC = V32_A_F | FO_IN_CONJ;
constexpr unsigned V32_A_M = V32_A_F | FO_IN_MASK;
constexpr unsigned V32_A_MP = V32_A_M | FO_PRE;
constexpr unsigned V32_A_MPC = V32_A_MP | FO_PRE_CONJ;
constexpr unsigned V32_B_F = FO_STORE_T | FO_OUT_MASK;                           // second pass of a plain transform
constexpr unsigned V32_B_FC = V32_B_F | FO_OUT_CONJ;
constexpr unsigned V32_B_N = FO_STORE_T;                                         // ... all rows kept: no store mask
constexpr unsigned V32_B_NC = V32_B_N | FO_OUT_CONJ;
constexpr unsigned V32_BM = FO_TWO_FFTS | FO_TWIDDLE;                            // middle pass: contiguous lines, in place
constexpr unsigned V32_BMC = V32_BM | FO_MID_CONJ;
constexpr unsigned V32_BM_N = FO_TWO_FFTS;                                       // ... its output twiddle left to the last pass
constexpr unsigned V32_BMC_N = V32_BM_N | FO_MID_CONJ;
constexpr unsigned V32_C_M = FO_LOAD_T | FO_STORE_T | FO_OUT_CONJ | FO_OUT_MASK; // last pass of a convolution
constexpr unsigned V32_C_N = FO_LOAD_T | FO_STORE_T | FO_OUT_CONJ;               // ... all rows kept: no store mask
constexpr unsigned V32_C_MP = V32_C_M | FO_POST;
constexpr unsigned V32_C_MPC = V32_C_MP | FO_POST_CONJ;
constexpr unsigned V32_C_TW = FO_IN_TWIDDLE;                                     // or-ed to V32_C_*: four-step twiddle on the loads
constexpr unsigned V32_A_H = V32_A_F | FO_IN_HALF;                               // Toeplitz with n = m = L/2: rows >= L/2 are padding ...
constexpr unsigned V32_C_H = V32_C_N | FO_OUT_HALF;                              // ... and rows >= L/2 of the result are dropped
constexpr unsigned V32_K_A = FO_LOAD_T | FO_STORE_T | FO_OUT_MASK;               // Kron: over i1 (stride), natural order out
constexpr unsigned V32_K_AC = V32_K_A | FO_IN_CONJ;
constexpr unsigned V32_K_B = FO_OUT_MASK;                                        // Kron: over i2 (contiguous)
constexpr unsigned V32_K_BC = FO_OUT_MASK | FO_OUT_CONJ;
// whole convolutions of FFT length 1024 in one kernel, a line = a column of the operand (fft_engine.cu: run_single_fast):
constexpr unsigned V32_1M = FO_TWO_FFTS | FO_OUT_CONJ | FO_DYN_LINES;            // Circulant(1024)
constexpr unsigned V32_1MC = V32_1M | FO_MID_CONJ;
constexpr unsigned V32_1H = V32_1M | FO_IN_HALF | FO_OUT_HALF;                   // Toeplitz 512 x 512: padding never loaded, dropped rows never stored
constexpr unsigned V32_1HC = V32_1H | FO_MID_CONJ;
constexpr unsigned V32_1K = V32_1M | FO_IN_MASK | FO_OUT_MASK;                   // other Toeplitz shapes padded to 1024: masks
constexpr unsigned V32_1KC = V32_1K | FO_MID_CONJ;

// `shape` selects the tile / occupancy instantiation: strided passes 0 = 8 lines, 2 CTAs per SM (128 registers), 1 = 8 lines,
// 3 CTAs per SM (80 registers); the middle pass of a convolution (warp-private lines, no CTA barrier) additionally
// 2 / 3 / 4 = 4 lines per CTA at 4 / 5 / 6 CTAs per SM (128 / 96 / 80 registers)
template <unsigned OPT, int LOGT, int MINB> int launch_v32_inst(const FastArgs<float2> &a, unsigned lines, cudaStream_t st) {
    constexpr bool LOAD_T = (OPT & FO_LOAD_T) != 0, STORE_T = (OPT & FO_STORE_T) != 0;
    // FMB_V32_SMEM_PAD (experiments): extra dynamic shared memory per CTA = a smaller L1 at unchanged occupancy.  These
    // passes read their twiddle tables through L1; ~27 KB of L1 instead of ~93 KB costs +24 % - which is what three CTAs
    // per SM (203 KB of shared memory) cost as well, whatever feeds them (profiles/r2_experiments.txt, calls 4, 19, 22, 23)
    static const size_t smem_pad = getenv("FMB_V32_SMEM_PAD") ? (size_t)atol(getenv("FMB_V32_SMEM_PAD")) : 0;
    const size_t smem = (size_t)(1 << LOGT) * V32_RS * sizeof(float2) + smem_pad;
    constexpr bool DYN = (OPT & FO_DYN_LINES) != 0;
    static_assert(!DYN || (!LOAD_T && !STORE_T), "a run-time line stride exists on the contiguous sides only");
    if (a.in_fs != (LOAD_T ? 1024 : 1) || (!DYN && a.in_is != (LOAD_T ? 1 : 1024)) || a.out_ks != (STORE_T ? 1024 : 1) ||
        (!DYN && a.out_is != (STORE_T ? 1 : 1024)) || a.I != 1024) {
        set_error("V32 pass: arguments do not describe the fixed 1024 x 1024 geometry");
        return FMB_ERR_VALUE;
    }
    static int attr_done = 0;
    if (!attr_done) {
        FMB_CUDA_OK(cudaFuncSetAttribute(v32_pass_kernel<OPT, LOGT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (const char *cv = getenv("FMB_V32_CARVEOUT"))          // experiments: shared-memory share of the unified L1 (percent)
            FMB_CUDA_OK(cudaFuncSetAttribute(v32_pass_kernel<OPT, LOGT, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(cv)));
        attr_done = 1;
    }
#ifdef V32_TIMING
    struct TimingDump {
        unsigned long long *buf = nullptr; unsigned *cnt = nullptr;
        void dump() {
            if (!buf) return;
            cudaDeviceSynchronize();
            unsigned n = 0;
            cudaMemcpy(&n, cnt, 4, cudaMemcpyDeviceToHost);
            n = n < (unsigned)V32_TREC ? n : (unsigned)V32_TREC;
            std::vector<unsigned long long> h((size_t)n * V32_TFIELDS);
            cudaMemcpy(h.data(), buf, h.size() * 8, cudaMemcpyDeviceToHost);
            double acc[V32_TFIELDS] = {};
            for (unsigned r = 0; r < n; ++r)
                for (int f = 1; f < V32_TFIELDS; ++f)
                    if (h[(size_t)r * V32_TFIELDS + f]) acc[f] += (double)(h[(size_t)r * V32_TFIELDS + f] - h[(size_t)r * V32_TFIELDS + f - 1]);
            fprintf(stderr, "V32_TIMING opt %u logt %d minb %d: %u warp records, mean cycles per phase:", OPT, LOGT, MINB, n);
            for (int f = 1; f < V32_TFIELDS; ++f) fprintf(stderr, " %.0f", n ? acc[f] / n : 0.0);
            fprintf(stderr, "\n");
        }
    };
    static TimingDump td;
    if (!td.buf) {
        debug_dumpers().push_back([]() { td.dump(); });
        FMB_CUDA_OK(cudaMalloc(&td.buf, (size_t)V32_TREC * V32_TFIELDS * 8));
        FMB_CUDA_OK(cudaMalloc(&td.cnt, 4));
        FMB_CUDA_OK(cudaMemset(td.cnt, 0, 4));
    }
    v32_pass_kernel<OPT, LOGT, MINB><<<lines >> LOGT, 32 << LOGT, smem, st>>>(a, td.buf, td.cnt);
#else
    v32_pass_kernel<OPT, LOGT, MINB><<<lines >> LOGT, 32 << LOGT, smem, st>>>(a);
#endif
    FMB_LAUNCH_OK();
    return FMB_OK;
}

template <unsigned OPT> int launch_v32_variant(const FastArgs<float2> &a, unsigned lines, int shape, cudaStream_t st) {
    if constexpr ((OPT & FO_TWO_FFTS) != 0) {
        switch (shape) {
#ifndef V32_LEAN
            case 1: return launch_v32_inst<OPT, 3, 3>(a, lines, st);
            case 2: return launch_v32_inst<OPT, 2, 4>(a, lines, st);
            case 3: return launch_v32_inst<OPT, 2, 5>(a, lines, st);
            case 4: return launch_v32_inst<OPT, 2, 6>(a, lines, st);
#endif
            default: break;
        }
        return launch_v32_inst<OPT, 3, 2>(a, lines, st);
    } else {
#ifndef V32_LEAN
        if (shape == 1) return launch_v32_inst<OPT, 3, 3>(a, lines, st);
#endif
        return launch_v32_inst<OPT, 3, 2>(a, lines, st);
    }
}

}  // namespace fmb
