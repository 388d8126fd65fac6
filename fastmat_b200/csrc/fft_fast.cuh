// Fast path of the FFT / convolution engine: power-of-two transforms in fastmat's column-major layout.
//
// Same algorithm as fft_pass.cuh (four-step FFT over an L2-resident intermediate; Stockham radix-16 stages staged in
// shared memory) but with the pass length R, the tile width T, the radix plan and every fused option fixed at compile
// time, so that all shared-memory offsets are immediates, there is no per-element flag test, and a thread's sixteen
// values stay in registers from the global load to the first exchange and from the last exchange to the global store.
//
// Thread t of a CTA owns, in every stage, the sixteen transform positions  jb + (R/16) m, m = 0..15  of one line
// (jb = its "row", the line = its "column"): a radix-P stage treats them as 16/P butterflies.  Two thread orders
// exist: line-fastest (consecutive threads walk the T lines of the tile: used whenever the lines are the contiguous
// direction in global memory) and row-fastest (consecutive threads walk jb: used when the transform axis is
// contiguous).  Shared memory is laid out [line][position] with one pad element per 16 positions and a line stride
// of 4 (mod 16) elements, which makes both orders bank-conflict free (see DESIGN.md "shared-memory layout").
#pragma once
#include "cx.cuh"

namespace fmb {

template <int LOGR> struct FastPlan;      // radix plan: first stage radix 16 (Ns = 1), then 16s, then the remainder
template <> struct FastPlan<6>  { static constexpr int S = 2; static constexpr int P0 = 16, P1 = 4,  P2 = 1; };
template <> struct FastPlan<7>  { static constexpr int S = 2; static constexpr int P0 = 16, P1 = 8,  P2 = 1; };
template <> struct FastPlan<8>  { static constexpr int S = 2; static constexpr int P0 = 16, P1 = 16, P2 = 1; };
template <> struct FastPlan<9>  { static constexpr int S = 3; static constexpr int P0 = 16, P1 = 16, P2 = 2; };
template <> struct FastPlan<10> { static constexpr int S = 3; static constexpr int P0 = 16, P1 = 16, P2 = 4; };
template <> struct FastPlan<11> { static constexpr int S = 3; static constexpr int P0 = 16, P1 = 16, P2 = 8; };
template <> struct FastPlan<12> { static constexpr int S = 3; static constexpr int P0 = 16, P1 = 16, P2 = 16; };

template <int LOGR> struct FastTile {     // lines per tile (log2): 4096 or 8192 elements per CTA
#ifndef FMB_FAST_TILE_LOG2
#define FMB_FAST_TILE_LOG2 13
#endif
    static constexpr int LOGT = (LOGR <= 9) ? (12 - LOGR) : (FMB_FAST_TILE_LOG2 - LOGR);
};

template <int LOGR, int LOGT> struct FastGeom {
    static constexpr int R = 1 << LOGR, T = 1 << LOGT, ROWS = R / 16, NT = ROWS * T;
    // line stride in shared memory: R positions + one pad per 16, then padded so that RS == 16/T (mod 16): a 64-bit
    // access is served per half-warp, and with that residue the 16 lanes of a half-warp (T lines x 16/T rows in
    // line-fastest order) fall on 16 different bank pairs
    static constexpr int WANT = (16 >> LOGT) ? (16 >> LOGT) : 1;
    static constexpr int RS = R + R / 16 + ((WANT - (R + R / 16)) % 16 + 16) % 16;
    static constexpr int SMEM_ELEMS = T * RS;
};

// runtime arguments of one fast pass (everything structural is a template parameter)
template <typename C> struct FastArgs {
    typedef typename real_of<C>::type S;
    const C *in;                 // element (f, line) at in[col*in_cs + f*in_fs + i*in_is]
    C *out;                      // element (k, line) at out[col*out_cs + k*out_ks + i*out_is]
    long long in_cs, out_cs;     // column strides (elements)
    int in_fs, in_is, out_ks, out_is;
    int I;                       // lines per column (power of two), log2 in logI
    int logI;
    int ncols;
    int in_n, in_lf, in_li;      // load mask: logical row f*in_lf + i*in_li must be < in_n
    int out_n, out_lk, out_li;   // store mask
    const C *tw;                 // stage twiddles of this pass length, see fast_stage_table_size()
    const C *twL, *twH;          // four-step twiddle W_N^e = twL[e & mask] * twH[e >> shift]
    int tw_shift; unsigned tw_mask;
    const C *twS;                // twS[i] = W_N^{(R/16) i}
    const C *mid;                // spectrum, transposed at plan creation: element (k, i) at mid[i*mid_is + k]
    int mid_is;
    const C *pre, *post;         // indexed by logical row
    const C *pf;                 // V32 first pass: same tile position in the slab whose input should be pulled into L2 now (or null)
};

// option bits of a pass
enum : unsigned {
    FO_LOAD_T = 1u,        // first stage line-fastest (else row-fastest)
    FO_STORE_T = 2u,       // last stage line-fastest
    FO_TWIDDLE = 4u,       // multiply the output by W_N^{i k}
    FO_TWO_FFTS = 8u,      // FFT -> * mid -> conj -> FFT
    FO_MID_CONJ = 16u,     // use conj(mid)
    FO_IN_CONJ = 32u,
    FO_OUT_CONJ = 64u,
    FO_IN_MASK = 128u,
    FO_OUT_MASK = 256u,
    FO_PRE = 512u, FO_PRE_CONJ = 1024u,
    FO_POST = 2048u, FO_POST_CONJ = 4096u,
    FO_IN_CG = 8192u,      // the input was written by other SMs in this launch (read through L2 only)
    FO_IN_HALF = 32768u,     // V32 passes: rows >= L/2 of the input are zero padding (Toeplitz): inputs m >= 16 of a thread are not loaded
    FO_OUT_HALF = 65536u,    // V32 passes: only rows < L/2 of the output are kept: outputs q >= 16 of a thread are not computed
    FO_IN_TWIDDLE = 16384u,  // V32 passes: multiply the INPUT by W_N^{i f} (the four-step twiddle moved out of the previous pass)
    FO_DYN_LINES = 131072u,  // V32 passes on contiguous lines: the line stride comes from in_is / out_is (else the fixed 1024)
};

template <typename C> __device__ __forceinline__ C ld_cg(const C *p) { return __ldcg(p); }
// streaming load of a pass input: each element is used once per launch and should not displace the twiddle tables from L1
__device__ __forceinline__ float2 ld_stream(const float2 *p) {
    float2 r;
    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ld_stream(const double2 *p) {
    double2 r;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

// read-only load whose position in the instruction stream is kept relative to the other loads of its kind (volatile,
// but no memory clobber): the spectrum loads of the middle pass are software-pipelined by hand, four values ahead of
// the butterfly that consumes them, instead of being all hoisted (register pressure) or all exposed (latency)
__device__ __forceinline__ float2 ld_nc_ordered(const float2 *p) {
    float2 r;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ld_nc_ordered(const double2 *p) {
    double2 r;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

// barrier between two phases of a pass.  GROUP: both phases use the row-fastest thread order, in which one line is
// owned by ROWS consecutive threads in every stage, so only those threads have to meet (a warp-level sync for short
// lines, a named barrier per line otherwise); the lines of a tile then drift apart and hide each other's latencies.
template <int ROWS, bool GROUP> __device__ __forceinline__ void fast_sync(int t) {
    if constexpr (!GROUP) __syncthreads();
    else if constexpr (ROWS <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(t + 1), "n"(ROWS) : "memory");
}

// Stage twiddles.  A radix-P stage with Ns finished sub-transforms multiplies input r of butterfly j by
// W_{Ns P}^{(j mod Ns) r}.  They are tabulated per stage as pairs (r = 2p, 2p+1) with j mod Ns as the fastest index:
//      table[p * Ns + kk] = { W^{kk 2p}, W^{kk (2p+1)} },   p < P/2, kk < Ns
// so that a thread fetches two twiddles per 128-bit load and the lanes of a warp (consecutive kk in either thread
// order) read consecutive addresses.  One pass length R has the tables of its second and third stage back to back.
template <typename C> struct alignas(2 * sizeof(C)) CPair { C a, b; };
__device__ __forceinline__ CPair<float2> ldg_pair(const CPair<float2> *p) {
    const float4 q = __ldg(reinterpret_cast<const float4 *>(p));
    CPair<float2> r; r.a = make_float2(q.x, q.y); r.b = make_float2(q.z, q.w);
    return r;
}
__device__ __forceinline__ CPair<double2> ldg_pair(const CPair<double2> *p) {
    CPair<double2> r;
    r.a = __ldg(reinterpret_cast<const double2 *>(p));
    r.b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return r;
}
template <int LOGR> constexpr int fast_stage_table_offset(int stage) {        // in CPair units; stage 1 or 2
    return stage == 1 ? 0 : (FastPlan<LOGR>::P1 / 2) * FastPlan<LOGR>::P0;
}

// one radix-P stage on the sixteen register values of a thread.
//   v[m] holds position jb + ROWS*m on entry (natural Stockham input order of every stage) and, on exit, the values
//   are written through `store(k, value)` at their Stockham output positions.  `tab`: this stage's twiddle table.
template <typename C, int LOGR, int P, int NS, typename Store>
__device__ __forceinline__ void fast_butterflies(C (&v)[16], int jb, const CPair<C> *__restrict__ tab, Store store) {
    constexpr int R = 1 << LOGR, ROWS = R / 16, NBF = 16 / P;
#pragma unroll
    for (int it = 0; it < NBF; ++it) {
        const int j = jb + it * ROWS;
        C u[P];
#pragma unroll
        for (int r = 0; r < P; ++r) u[r] = v[it + NBF * r];
        if constexpr (NS > 1) {
            // j mod Ns: the part contributed by `it` is a compile-time offset when a thread's butterflies stay inside Ns
            const CPair<C> *tp;
            if constexpr (NS > ROWS * (NBF - 1)) tp = tab + (jb & (NS - 1)) + ((it * ROWS) & (NS - 1));
            else tp = tab + (j & (NS - 1));
#pragma unroll
            for (int p2 = 0; p2 < P / 2; ++p2) {
                const CPair<C> w = ldg_pair(tp + p2 * NS);
                if (p2 > 0) u[2 * p2] = cmul(u[2 * p2], w.a);
                u[2 * p2 + 1] = cmul(u[2 * p2 + 1], w.b);
            }
        }
        if constexpr (P == 2) dft2(u[0], u[1]);
        else if constexpr (P == 4) dft4(u[0], u[1], u[2], u[3]);
        else if constexpr (P == 8) dft8(u);
        else dft16(u);
        const int j0 = (NS * P == R) ? j : (((j & ~(NS - 1)) * P) | (j & (NS - 1)));   // last stage: j < NS
#pragma unroll
        for (int q = 0; q < P; ++q) store(j0 + q * NS, u[outpos<P>(q)], it, q);
    }
}

template <int P_, int NS_> struct FastTag { static constexpr int P = P_, NS = NS_; };

template <int LOGR, int LOGT, bool ORDER_T> __device__ __forceinline__ void fast_thread_pos(int tid, int &jb, int &t) {
    constexpr int ROWS = (1 << LOGR) / 16, T = 1 << LOGT;
    if (ORDER_T) { t = tid & (T - 1); jb = tid >> LOGT; }
    else { jb = tid & (ROWS - 1); t = tid >> (LOGR - 4); }
}

// The whole pass for one tile.  `tile` indexes T consecutive lines; line = col*I + i.
// PART: 0 = the whole pass; for two-transform passes 1 = up to the spectrum multiply, 2 = the second transform.  The two
// halves are compiled as separate (non-inlined) functions so that each gets its own 64-register allocation.
template <typename C, int LOGR, int LOGT, unsigned OPT, int PART>
__device__ __forceinline__ void fast_pass_part(const FastArgs<C> &a, unsigned tile, C *smem, long long in_off = 0,
                                               long long out_off = 0) {
    typedef FastGeom<LOGR, LOGT> G;
    typedef FastPlan<LOGR> PL;
    constexpr int R = G::R, ROWS = G::ROWS, RS = G::RS;
    constexpr bool LOAD_T = (OPT & FO_LOAD_T) != 0, STORE_T = (OPT & FO_STORE_T) != 0, TWO = (OPT & FO_TWO_FFTS) != 0;
    constexpr int NS1 = PL::P0, NS2 = PL::P0 * PL::P1;
    const int tid = threadIdx.x;
    const unsigned line0 = tile << LOGT;
    const unsigned col = line0 >> a.logI;                  // T divides I: a tile never straddles two columns
    const unsigned i0 = line0 & (unsigned)(a.I - 1);
    C v[16];
    int jb, t;
    const CPair<C> *const tab1 = reinterpret_cast<const CPair<C> *>(a.tw);
    const CPair<C> *const tab2 = tab1 + fast_stage_table_offset<LOGR>(2);

    // ------------------------------------------------------------------ stage 0: global -> registers -> shared
    fast_thread_pos<LOGR, LOGT, LOAD_T>(tid, jb, t);
    if constexpr (PART != 2) {
        const unsigned i = i0 + t;
        const C *src = a.in + in_off + (long long)col * a.in_cs + (long long)i * a.in_is + (long long)jb * a.in_fs;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int f = jb + ROWS * m;
            bool ok = true;
            if (OPT & FO_IN_MASK) ok = (f * a.in_lf + (int)i * a.in_li) < a.in_n;
            C val = mk<C>(0, 0);
            if (ok) {
                const C *p = src + (long long)(ROWS * m) * a.in_fs;
                val = (OPT & FO_IN_CG) ? ld_cg(p) : ld_stream(p);
                if (OPT & FO_IN_CONJ) val = cconj(val);
                if (OPT & FO_PRE) {
                    const C w = __ldg(a.pre + (f * a.in_lf + (int)i * a.in_li));
                    val = (OPT & FO_PRE_CONJ) ? cmulc(val, w) : cmul(val, w);
                }
            }
            v[m] = val;
        }
        C *sline = smem + t * RS;
        fast_butterflies<C, LOGR, PL::P0, 1>(v, jb, tab1, [&](int k, C val, int, int) { sline[k + (k >> 4)] = val; });
    }

    // inner stages (shared -> shared) always use the row-fastest order: conflict free, and their barriers are per line
#ifndef FMB_FAST_INNER_T
#define FMB_FAST_INNER_T 0
#endif
    constexpr bool INNER_T = FMB_FAST_INNER_T != 0;
    constexpr bool G0 = !LOAD_T, GL = !STORE_T, GI = !INNER_T;    // is the first / last / an inner stage row-fastest?
    // ------------------------------------------------------------------ stage 1 (and 2): shared -> shared, or the last stage
    auto load16 = [&](const C *sl, int jb_) {
        const C *b = sl + jb_ + (jb_ >> 4);
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = b[ROWS * m + (ROWS * m >> 4)];      // ROWS*m is a multiple of 16 when ROWS >= 16
    };
    static_assert(ROWS % 16 == 0 || ROWS == 8 || ROWS == 4, "row count must keep the pad offsets separable");
    auto load16g = [&](const C *sl, int jb_) {
        if constexpr (ROWS % 16 == 0) load16(sl, jb_);
        else {
#pragma unroll
            for (int m = 0; m < 16; ++m) { const int f = jb_ + ROWS * m; v[m] = sl[f + (f >> 4)]; }
        }
    };

    // generic "last stage + global store" used by both the single and the second transform
    auto final_store = [&](int jb_, int t_, auto stage_tag) {
        constexpr int P = decltype(stage_tag)::P, NS = decltype(stage_tag)::NS;
        const unsigned i = i0 + t_;
        C *dst = a.out + out_off + (long long)col * a.out_cs + (long long)i * a.out_is;
        // four-step twiddle of output k = jb + ROWS*(it + NBF*q):  W^{i jb} * (W^{ROWS i})^{it} * ((W^{ROWS i})^{NBF})^q
        constexpr int NBF = 16 / P;
        C wbase = mk<C>(1, 0), s1 = mk<C>(1, 0), sq = mk<C>(1, 0), wcur = mk<C>(1, 0);
        if (OPT & FO_TWIDDLE) {
            const unsigned e = i * (unsigned)jb_;
            wbase = cmul(__ldg(a.twL + (e & a.tw_mask)), __ldg(a.twH + (e >> a.tw_shift)));
            s1 = __ldg(a.twS + i);
            sq = s1;                                  // sq = s1^NBF by repeated squaring (NBF is a power of two)
#pragma unroll
            for (int b2 = 1; b2 < NBF; b2 *= 2) sq = cmul(sq, sq);
        }
        fast_butterflies<C, LOGR, P, NS>(v, jb_, (NS == NS1 ? tab1 : tab2), [&](int k, C val, int it, int q) {
            if (OPT & FO_TWIDDLE) {
                if (q == 0) {                          // start of butterfly `it`: wcur = wbase * s1^it
                    if (it > 0) wbase = cmul(wbase, s1);
                    wcur = wbase;
                } else {
                    wcur = cmul(wcur, sq);
                }
                val = cmul(val, wcur);
            }
            if (OPT & FO_OUT_CONJ) val = cconj(val);
            bool ok = true;
            const int mrow = k * a.out_lk + (int)i * a.out_li;
            if (OPT & FO_OUT_MASK) ok = mrow < a.out_n;
            if (OPT & FO_POST) {
                if (ok) {
                    const C pw = __ldg(a.post + mrow);
                    val = (OPT & FO_POST_CONJ) ? cmulc(val, pw) : cmul(val, pw);
                }
            }
            if (ok) dst[(long long)k * a.out_ks] = val;
        });
    };

    typedef FastTag<PL::P1, NS1> Tag1;
    typedef FastTag<PL::P2, NS2> Tag2;

    auto smem_store = [&](C *sl) { return [sl](int k, C val, int, int) { sl[k + (k >> 4)] = val; }; };

    if constexpr (!TWO) {
        if constexpr (PL::S == 2) {
            fast_sync<ROWS, G0 && GL>(t);
            fast_thread_pos<LOGR, LOGT, STORE_T>(tid, jb, t);
            load16g(smem + t * RS, jb);
            final_store(jb, t, Tag1());
        } else {
            fast_sync<ROWS, G0 && GI>(t);
            fast_thread_pos<LOGR, LOGT, INNER_T>(tid, jb, t);
            load16g(smem + t * RS, jb);
            fast_sync<ROWS, GI>(t);
            fast_butterflies<C, LOGR, PL::P1, NS1>(v, jb, tab1, smem_store(smem + t * RS));
            fast_sync<ROWS, GL && GI>(t);
            fast_thread_pos<LOGR, LOGT, STORE_T>(tid, jb, t);
            load16g(smem + t * RS, jb);
            final_store(jb, t, Tag2());
        }
    } else {
        // ---- finish the first transform in shared memory; its last stage multiplies every output X[k] by the spectrum
        //      (and conjugates: the second transform runs the inverse as conj(FFT(conj(.))))
        if constexpr (PART != 2) fast_sync<ROWS, G0 && GI>(t);
        fast_thread_pos<LOGR, LOGT, INNER_T>(tid, jb, t);
        if constexpr (PART != 2) {
        // spectrum values of this thread's sixteen outputs k = jb + ROWS*n', loaded four at a time one batch ahead
        const C *mp = a.mid + (long long)(i0 + t) * a.mid_is + jb;
        C mv[2][4];
        auto mid_issue = [&](int b, auto stage_tag) {
            constexpr int P = decltype(stage_tag)::P, NBF = 16 / P;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int n = 4 * b + e, it = n / P, q = n % P;
                mv[b & 1][e] = ld_nc_ordered(mp + ROWS * (it + NBF * q));
            }
        };
        auto mid_store = [&](C *sl, auto stage_tag) {
            return [sl, &mv, &mid_issue, stage_tag](int k, C val, int it, int q) {
                constexpr int P = decltype(stage_tag)::P;
                const int n = it * P + q;
                if (n % 4 == 0 && n + 4 < 16) mid_issue(n / 4 + 1, stage_tag);
                const C w = mv[(n / 4) & 1][n % 4];
                sl[k + (k >> 4)] = cconj((OPT & FO_MID_CONJ) ? cmulc(val, w) : cmul(val, w));
            };
        };
        if constexpr (PL::S == 2) {
            mid_issue(0, Tag1());
            load16g(smem + t * RS, jb);
            fast_sync<ROWS, GI>(t);
            fast_butterflies<C, LOGR, PL::P1, NS1>(v, jb, tab1, mid_store(smem + t * RS, Tag1()));
            fast_sync<ROWS, GI>(t);
        } else {
            load16g(smem + t * RS, jb);
            fast_sync<ROWS, GI>(t);
            fast_butterflies<C, LOGR, PL::P1, NS1>(v, jb, tab1, smem_store(smem + t * RS));
            fast_sync<ROWS, GI>(t);
            mid_issue(0, Tag2());
            load16g(smem + t * RS, jb);
            fast_sync<ROWS, GI>(t);
            fast_butterflies<C, LOGR, PL::P2, NS2>(v, jb, tab2, mid_store(smem + t * RS, Tag2()));
            fast_sync<ROWS, GI>(t);
        }
        }
        if constexpr (PART != 1) {
        // ---- second transform
        load16g(smem + t * RS, jb);
        fast_sync<ROWS, GI>(t);
        fast_butterflies<C, LOGR, PL::P0, 1>(v, jb, tab1, smem_store(smem + t * RS));
        if constexpr (PL::S == 2) {
            fast_sync<ROWS, GL && GI>(t);
            fast_thread_pos<LOGR, LOGT, STORE_T>(tid, jb, t);
            load16g(smem + t * RS, jb);
            final_store(jb, t, Tag1());
        } else {
            fast_sync<ROWS, GI>(t);
            load16g(smem + t * RS, jb);
            fast_sync<ROWS, GI>(t);
            fast_butterflies<C, LOGR, PL::P1, NS1>(v, jb, tab1, smem_store(smem + t * RS));
            fast_sync<ROWS, GL && GI>(t);
            fast_thread_pos<LOGR, LOGT, STORE_T>(tid, jb, t);
            load16g(smem + t * RS, jb);
            final_store(jb, t, Tag2());
        }
        }
    }
}

template <typename C, int LOGR, int LOGT, unsigned OPT, int PART>
__device__ __noinline__ void fast_pass_half(const FastArgs<C> &a, unsigned tile, long long in_off = 0, long long out_off = 0) {
    extern __shared__ __align__(16) unsigned char fmb_fast_smem[];
    fast_pass_part<C, LOGR, LOGT, OPT, PART>(a, tile, reinterpret_cast<C *>(fmb_fast_smem), in_off, out_off);
}

// a whole pass as non-inlined function(s): used by the fused persistent kernel, where several pass types share one kernel
template <typename C, int LOGR, int LOGT, unsigned OPT>
__device__ __forceinline__ void fast_pass_call(const FastArgs<C> &a, unsigned tile, long long in_off, long long out_off) {
    if constexpr (OPT & FO_TWO_FFTS) {
        fast_pass_half<C, LOGR, LOGT, OPT, 1>(a, tile, in_off, out_off);
        fast_pass_half<C, LOGR, LOGT, OPT, 2>(a, tile, in_off, out_off);
    } else {
        fast_pass_half<C, LOGR, LOGT, OPT, 0>(a, tile, in_off, out_off);
    }
}

template <typename C, int LOGR, int LOGT, unsigned OPT>
__device__ __forceinline__ void fast_pass_tile(const FastArgs<C> &a, unsigned tile, C *smem) {
    if constexpr (OPT & FO_TWO_FFTS) {
        fast_pass_half<C, LOGR, LOGT, OPT, 1>(a, tile);
        fast_pass_half<C, LOGR, LOGT, OPT, 2>(a, tile);
    } else {
        fast_pass_part<C, LOGR, LOGT, OPT, 0>(a, tile, smem);
    }
}

template <typename C, int LOGR, int LOGT, unsigned OPT>
__global__ void __launch_bounds__(FastGeom<LOGR, LOGT>::NT, (FastGeom<LOGR, LOGT>::NT >= 512 ? 2 : 3))
fast_pass_kernel(const __grid_constant__ FastArgs<C> a) {
    extern __shared__ __align__(16) unsigned char fmb_fast_smem[];
    fast_pass_tile<C, LOGR, LOGT, OPT>(a, blockIdx.x, reinterpret_cast<C *>(fmb_fast_smem));
}

}  // namespace fmb
