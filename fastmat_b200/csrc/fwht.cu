// Fast Walsh-Hadamard transform (natural / Sylvester order, unnormalised) for all eight fastmat dtypes.
//
// Reference: fastmat/Hadamard.pyx:164-230 (_forwardC) applies, per column, `order` in-place radix-2 sweeps with
// butterfly distance 1, 2, 4, ... and the butterfly (a, b) -> (a + b, a - b) of _hadamardCore (:36-61) in the
// array's own dtype.  Here a column is processed in one or more shared-memory passes; pass p covers a contiguous
// range of index bits [s, s+b) and inside a pass the bits are taken in ASCENDING order, so the sequence of
// additions every output element sees is exactly the reference's: integer results are bit-exact by ring
// arithmetic, and floating-point results are bit-exact too because the rounding sequence is identical.
//
// Per pass one CTA owns a tile [2^b "mid" values] x [T lines]; a line is a (column, low-bits) pair for the
// column-major layout, or a column for the row-major layout, so that the contiguous memory direction is always
// the one consecutive threads walk.  Passes after the first run in place on the output; the host drives them
// slab-of-columns by slab so that the slab is still L2-resident when the next pass reads it.
#include "common.h"
#include "cx.cuh"

namespace fmb {

template <typename T> struct HOps {
    static FMB_HD T add(T a, T b) { return a + b; }
    static FMB_HD T sub(T a, T b) { return a - b; }
};
template <> struct HOps<int8_t> {
    static FMB_HD int8_t add(int8_t a, int8_t b) { return (int8_t)(uint8_t)((unsigned)(uint8_t)a + (unsigned)(uint8_t)b); }
    static FMB_HD int8_t sub(int8_t a, int8_t b) { return (int8_t)(uint8_t)((unsigned)(uint8_t)a - (unsigned)(uint8_t)b); }
};
template <> struct HOps<int16_t> {
    static FMB_HD int16_t add(int16_t a, int16_t b) { return (int16_t)(uint16_t)((unsigned)(uint16_t)a + (unsigned)(uint16_t)b); }
    static FMB_HD int16_t sub(int16_t a, int16_t b) { return (int16_t)(uint16_t)((unsigned)(uint16_t)a - (unsigned)(uint16_t)b); }
};
template <> struct HOps<int32_t> {
    static FMB_HD int32_t add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
    static FMB_HD int32_t sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
};
template <> struct HOps<int64_t> {
    static FMB_HD int64_t add(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
    static FMB_HD int64_t sub(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }
};
template <> struct HOps<float2> {
    static FMB_HD float2 add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
    static FMB_HD float2 sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
};
template <> struct HOps<double2> {
    static FMB_HD double2 add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
    static FMB_HD double2 sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
};

struct FwhtPass {
    int b;                    // bits handled by this pass
    int s;                    // lowest bit handled
    int T;                    // lines per tile
    int mid_contig;           // 1: the mid index is the contiguous memory direction (column-major, s == 0)
    int pad;                  // shared-memory padding: index e -> e + (e >> 4) when set
    long long lines_total;    // number of lines
    long long lo_count;       // 2^s (column-major) or 1 (row-major: lines are columns)
    long long hi_count;       // 2^(order - s - b)
    long long ncols;
    int row_major;            // line decode: 0: line = (c*hi_count + hi)*lo_count + lo ; 1: line = (hi*lo_count + lo)*ncols + c
    const void *in;
    void *out;
    long long in_rs, in_cs, out_rs, out_cs;
};

template <typename T> FMB_HD void fwht_line_base(const FwhtPass &p, long long line, long long &in_off, long long &out_off) {
    long long c, hi, lo;
    if (p.row_major) {
        long long q = line / p.ncols;
        c = line - q * p.ncols;
        hi = q / p.lo_count;
        lo = q - hi * p.lo_count;
    } else {
        lo = line % p.lo_count;
        long long q = line / p.lo_count;
        c = q / p.hi_count;
        hi = q - c * p.hi_count;
    }
    long long n = (hi << (p.s + p.b)) | lo;
    in_off = c * p.in_cs + n * p.in_rs;
    out_off = c * p.out_cs + n * p.out_rs;
}

template <typename T, int KK> FMB_HD void fwht_substage(T *sm, int u, int b, int T_, int pad, int tid, int NT) {
    constexpr int E = 1 << KK;
    const int groups = (1 << (b - KK)) * T_;
    for (int g = tid; g < groups; g += NT) {
        int l = g % T_;
        int mg = g / T_;
        int mid_base = ((mg >> u) << (u + KK)) | (mg & ((1 << u) - 1));
        T v[E];
#pragma unroll
        for (int r = 0; r < E; ++r) {
            int e = (mid_base + (r << u)) * T_ + l;
            v[r] = sm[pad ? e + (e >> 4) : e];
        }
#pragma unroll
        for (int lev = 0; lev < KK; ++lev) {          // ascending bit order inside the sub-stage
            const int d = 1 << lev;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                if ((r & d) == 0) {
                    T a = v[r], bb = v[r | d];
                    v[r] = HOps<T>::add(a, bb);
                    v[r | d] = HOps<T>::sub(a, bb);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < E; ++r) {
            int e = (mid_base + (r << u)) * T_ + l;
            sm[pad ? e + (e >> 4) : e] = v[r];
        }
    }
}

template <typename T, typename Sync>
FMB_HD void fwht_pass_body(const FwhtPass &p, long long tile, int tid, int NT, T *sm, Sync &sync) {
    const int nmid = 1 << p.b;
    const int telems = nmid * p.T;
    const long long step_in = ((long long)1 << p.s) * p.in_rs, step_out = ((long long)1 << p.s) * p.out_rs;
    // ---- load tile (memory order)
    for (int e = tid; e < telems; e += NT) {
        int mid, t;
        if (p.mid_contig) { t = e / nmid; mid = e - t * nmid; }
        else { mid = e / p.T; t = e - mid * p.T; }
        long long line = tile * p.T + t;
        if (line < p.lines_total) {
            long long io, oo;
            fwht_line_base<T>(p, line, io, oo);
            int se = mid * p.T + t;
            sm[p.pad ? se + (se >> 4) : se] = ((const T *)p.in)[io + (long long)mid * step_in];
        }
    }
    sync();
    // ---- butterflies, ascending bit order, up to 4 bits per sub-stage
    for (int u = 0; u < p.b;) {
        int kk = p.b - u;
        if (kk > 4) kk = 4;
        switch (kk) {
            case 4: fwht_substage<T, 4>(sm, u, p.b, p.T, p.pad, tid, NT); break;
            case 3: fwht_substage<T, 3>(sm, u, p.b, p.T, p.pad, tid, NT); break;
            case 2: fwht_substage<T, 2>(sm, u, p.b, p.T, p.pad, tid, NT); break;
            default: fwht_substage<T, 1>(sm, u, p.b, p.T, p.pad, tid, NT); break;
        }
        u += kk;
        sync();
    }
    // ---- store tile
    for (int e = tid; e < telems; e += NT) {
        int mid, t;
        if (p.mid_contig) { t = e / nmid; mid = e - t * nmid; }
        else { mid = e / p.T; t = e - mid * p.T; }
        long long line = tile * p.T + t;
        if (line < p.lines_total) {
            long long io, oo;
            fwht_line_base<T>(p, line, io, oo);
            int se = mid * p.T + t;
            ((T *)p.out)[oo + (long long)mid * step_out] = sm[p.pad ? se + (se >> 4) : se];
        }
    }
}

struct FwhtDevSync {
    FMB_HD void operator()() const {
#ifdef __CUDA_ARCH__
        __syncthreads();
#endif
    }
};

template <typename T> __global__ void __launch_bounds__(256) fwht_pass_kernel(const __grid_constant__ FwhtPass p) {
    extern __shared__ __align__(16) unsigned char fwht_smem_raw[];
    FwhtDevSync sync;
    fwht_pass_body<T, FwhtDevSync>(p, (long long)blockIdx.x, (int)threadIdx.x, (int)blockDim.x, reinterpret_cast<T *>(fwht_smem_raw), sync);
}

template <typename T> static int launch_fwht_pass(const FwhtPass &p, cudaStream_t st) {
    const int nmid = 1 << p.b;
    size_t elems = (size_t)nmid * p.T;
    size_t smem = (elems + (p.pad ? (elems >> 4) + 1 : 0)) * sizeof(T);
    long long tiles = (p.lines_total + p.T - 1) / p.T;
    if (tiles <= 0) return FMB_OK;
    if (tiles > 2147483647LL) { set_error("FWHT: too many tiles"); return FMB_ERR_VALUE; }
#ifdef FMB_EMULATE
    emulate_launch(tiles, 64, smem, [&](long long tile, int tid, int nt, void *sm, const std::function<void()> &bar) {
        struct S { const std::function<void()> &b; void operator()() const { b(); } } sync{bar};
        fwht_pass_body<T, S>(p, tile, tid, nt, (T *)sm, sync);
    });
    g_launches.fetch_add(1);
    return FMB_OK;
#else
    static int attr_done = 0;
    if (!attr_done) {
        FMB_CUDA_OK(cudaFuncSetAttribute(fwht_pass_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 << 10));
        attr_done = 1;
    }
    fwht_pass_kernel<T><<<(unsigned)tiles, 256, smem, st>>>(p);
    FMB_LAUNCH_OK();
    return FMB_OK;
#endif
}

// ------------------------------------------------------------------------------------------- fast path
// Column-major arrays, order >= 8: every thread keeps sixteen values in registers per sub-stage (four index bits, taken
// in ascending order so the result stays bit-identical to the reference), sub-stages exchange through padded shared
// memory, the first sub-stage reads global memory directly and the last one writes it directly.
struct FwhtFastPass {
    const void *in;
    void *out;
    long long in_cs, out_cs;   // column strides (elements); row stride is 1
    int b, s;                  // this pass transforms index bits [s, s+b)
    int logT;                  // lines (consecutive low-bit values) per tile; 0 for the contiguous first pass
    int order;
    long long tiles_per_col;   // a power of two ...
    int log_tpc;               // ... and its log2: tile -> (column, tile in column) by shift and mask, no 64-bit division
    unsigned long long pol_in, pol_out;   // L2 cache-hint policies of the loads / stores (4-byte types; 0: none)
};

// 4-byte accesses with an L2 eviction policy (createpolicy encodings: see fwht_fast)
template <typename T> __device__ __forceinline__ T fwht_ld_hint(const T *p, unsigned long long pol) {
    if constexpr (sizeof(T) == 4) {
        unsigned r;
        asm volatile("ld.global.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
        return reinterpret_cast<const T &>(r);
    } else return *p;
}
template <typename T> __device__ __forceinline__ void fwht_st_hint(T *p, T v, unsigned long long pol) {
    if constexpr (sizeof(T) == 4) {
        asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(p), "r"(reinterpret_cast<const unsigned &>(v)), "l"(pol) : "memory");
    } else *p = v;
}

template <typename T, int LEV0> __device__ __forceinline__ void fwht16(T (&v)[16]) {
#pragma unroll
    for (int lev = LEV0; lev < 4; ++lev) {              // levels below LEV0 were transformed by the previous sub-stage
        const int d = 1 << lev;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            if ((r & d) == 0) {
                const T a = v[r], b = v[r | d];
                v[r] = HOps<T>::add(a, b);
                v[r | d] = HOps<T>::sub(a, b);
            }
        }
    }
}

// sub-stages after the first one, resolved at compile time: the sub-stage starting at bit U covers bits [UU, UU+4) where
// UU = min(U, B-4) (fewer than four bits left: re-use the top four positions and skip the levels already done)
template <typename T, int B, int U>
__device__ __forceinline__ void fwht_sub_chain(T (&v)[16], T *sl, T *gout_t, int q, long long step, unsigned long long pol_out) {
    constexpr int UU = (U + 4 > B) ? (B - 4) : U;
    constexpr int LEV0 = U - UU;
    const int l = q & ((1 << UU) - 1), h = q >> UU;
    const int m0 = (h << (UU + 4)) | l;
    const T *base = sl + m0 + (m0 >> 4);
    if constexpr (UU >= 4) {
        // (r << UU) is a multiple of 16: the pad term separates -> immediate offsets
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = base[(r << UU) + ((r << UU) >> 4)];
    } else {
#pragma unroll
        for (int r = 0; r < 16; ++r) { const int m = m0 + (r << UU); v[r] = sl[m + (m >> 4)]; }
    }
    fwht16<T, LEV0>(v);
    if constexpr (UU + 4 >= B) {                        // last sub-stage: straight to global memory
        T *dst = gout_t + (long long)m0 * step;
        if (sizeof(T) == 4 && pol_out) {
#pragma unroll
            for (int r = 0; r < 16; ++r) fwht_st_hint(dst + (long long)(r << UU) * step, v[r], pol_out);
        } else {
#pragma unroll
            for (int r = 0; r < 16; ++r) dst[(long long)(r << UU) * step] = v[r];
        }
    } else {
        __syncthreads();
        T *wb = sl + m0 + (m0 >> 4);
#pragma unroll
        for (int r = 0; r < 16; ++r) wb[(r << UU) + ((r << UU) >> 4)] = v[r];
        __syncthreads();
        fwht_sub_chain<T, B, U + 4>(v, sl, gout_t, q, step, pol_out);
    }
}

// One tile of one pass.  A tile is all 2^B "mid" values x Tn lines; a line is a low-bit value (strided passes) or, for the
// contiguous first pass, one chunk of 2^B consecutive elements.  CG: read through L2 only (the input was
// written by other SMs earlier in the same launch).
template <typename T, int B, bool CONTIG, bool CG, int LOGSTEP = -1>
__device__ __forceinline__ void fwht_tile(const FwhtFastPass &p, long long tile, T *sm, int tid) {
    const int logT = p.logT, Tn = 1 << logT;
    const long long col = tile >> p.log_tpc;
    const long long tin = tile & (p.tiles_per_col - 1);
    const int t = tid & (Tn - 1), q = tid >> logT;      // line-fastest thread order
    long long base;
    if (CONTIG) {
        base = ((tin << logT) + t) << B;                // chunk (tin*Tn + t) of the column
    } else {
        // tile -> (hi, lo0): Tn consecutive lo values starting at lo0, one hi value
        const int log_lo = (LOGSTEP >= 0 ? LOGSTEP : p.s) - logT;                 // lo_tiles = 2^s / Tn
        const long long hi = tin >> log_lo, lo0 = (tin & (((long long)1 << log_lo) - 1)) << logT;
        base = (hi << (p.s + B)) + lo0 + t;
    }
    const T *gin = (const T *)p.in + col * p.in_cs + base;
    T *gout = (T *)p.out + col * p.out_cs + base;
    constexpr int RSL = (1 << B) + ((1 << B) >> 4) + 1; // padded line stride in shared memory (odd -> conflict free)
    T *sl = sm + t * RSL;
    // LOGSTEP >= 0: the stride of the strided pass is a compile-time constant (order 20: 2^12 elements), so the sixteen
    // loads and stores of a thread are one base register plus immediates instead of a 64-bit multiply-add each
    const long long step = CONTIG ? 1ll : (LOGSTEP >= 0 ? ((long long)1 << (LOGSTEP >= 0 ? LOGSTEP : 0)) : ((long long)1 << p.s));
    T v[16];
    // ---- first sub-stage: bits [0, 4) of mid straight from global memory
    const T *src = gin + (long long)(q << 4) * step;
    if (CONTIG && !CG && sizeof(T) == 4 && (reinterpret_cast<unsigned long long>(src) & 15ull) == 0) {   // 4 x 128-bit loads
        const float4 *s4 = reinterpret_cast<const float4 *>(src);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float4 f = s4[r];
            v[4 * r + 0] = reinterpret_cast<T &>(f.x); v[4 * r + 1] = reinterpret_cast<T &>(f.y);
            v[4 * r + 2] = reinterpret_cast<T &>(f.z); v[4 * r + 3] = reinterpret_cast<T &>(f.w);
        }
    } else {
        if (sizeof(T) == 4 && p.pol_in) {
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = fwht_ld_hint(src + (long long)r * step, p.pol_in);
        } else {
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = CG ? __ldcg(src + (long long)r * step) : src[(long long)r * step];
        }
    }
    fwht16<T, 0>(v);
    if constexpr (B == 4) {                             // a single sub-stage: back to global memory directly
        T *dst = gout + (long long)(q << 4) * step;
#pragma unroll
        for (int r = 0; r < 16; ++r) dst[(long long)r * step] = v[r];
    } else {
        T *wb = sl + 17 * q;                            // m = 16 q + r  ->  m + (m >> 4) = 17 q + r
#pragma unroll
        for (int r = 0; r < 16; ++r) wb[r] = v[r];
        __syncthreads();
        fwht_sub_chain<T, B, 4>(v, sl, gout, q, step, p.pol_out);
    }
}

template <typename T, int B, bool CONTIG, int LOGSTEP = -1> __global__ void __launch_bounds__(512) fwht_fast_kernel(const __grid_constant__ FwhtFastPass p) {
    extern __shared__ __align__(16) unsigned char fwht_smem_raw[];
    fwht_tile<T, B, CONTIG, false, LOGSTEP>(p, (long long)blockIdx.x, reinterpret_cast<T *>(fwht_smem_raw), (int)threadIdx.x);
}


// ------------------------------------------------------------------------------------------- first pass, 4-byte types
// Bits [0, 12) of a column-major column in TWO register sub-stages of six bits (64 values per thread, one exchange)
// instead of three of four bits (two exchanges): a third less traffic through the load/store pipe, which - not HBM - is
// what bounds the transform once the intermediate lives in L2 (profiles/r2_hadamard_o20_f32_ncu_full.txt: mio_throttle).
//   * one CTA of 64 threads owns one line of 4096 contiguous elements (16 KB); each warp pulls its 8 KB half into shared
//     memory with one bulk copy (cp.async.bulk, completion on an mbarrier): nothing passes through the load/store pipe or
//     the registers on the way in, and the copy needs no coalescing from the thread layout;
//   * sub-stage 1, bits [0, 6): thread u owns the 64 consecutive elements 64 u ... 64 u + 63, i.e. sixteen 16-byte chunks
//     of its own 256-byte row.  All threads reading chunk s of their row at once would hit the same four banks, so lane k
//     (= u mod 8) reads chunk s ^ k into register slot s (conflict free: the eight lanes of a quarter warp touch eight
//     different chunks).  Slot s then holds a lane-dependent chunk, but slots s and s ^ d still hold partner chunks, and
//     which of the two is the "low" one only decides where the sum and where the difference go:
//         slot s <- v[s ^ d] + sigma v[s],  slot s ^ d <- v[s] - sigma v[s ^ d],   sigma = +1 / -1 by bit d of k
//     - one fused multiply-add each, exact (multiplying by +-1 does not round), so every output is still produced by the
//     reference's sequence of additions, bit for bit.  The row is written back in place the same skewed way;
//   * sub-stage 2, bits [6, 12): thread u owns elements u + 64 j (lanes read consecutive words: conflict free), six plain
//     levels, and 64 coalesced 128-byte stores per warp straight to global memory.
template <typename T> struct HPm;          // b + sigma * a  for sigma = +-1
template <> struct HPm<float> {
    typedef float S;
    static __device__ __forceinline__ float pm(float sigma, float a, float b) { return __fmaf_rn(sigma, a, b); }
};
template <> struct HPm<int32_t> {
    typedef int32_t S;
    static __device__ __forceinline__ int32_t pm(int32_t sigma, int32_t a, int32_t b) { return (int32_t)((uint32_t)sigma * (uint32_t)a + (uint32_t)b); }
};

__device__ __forceinline__ unsigned fwht_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// lane 0 of a warp: request this warp's 8 KB half of the line into `half` (completion on `bar`)
template <typename T> __device__ __forceinline__ void fwht12_issue(const T *gsrc, unsigned half_s, unsigned bar, unsigned long long pol_in) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(8192u) : "memory");
    if (pol_in)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(half_s), "l"(gsrc), "r"(8192u), "r"(bar), "l"(pol_in) : "memory");
    else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(half_s), "l"(gsrc), "r"(8192u), "r"(bar) : "memory");
}
__device__ __forceinline__ void fwht_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "FWHT12_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra FWHT12_DONE_%=;\n\t"
        "bra FWHT12_WAIT_%=;\n\t"
        "FWHT12_DONE_%=:\n\t}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}

// the two register sub-stages of one line by 64 threads (tid64): `buf` holds the line (this thread's warp has waited for
// its half), `group_sync` is a barrier of the 64 threads
template <typename T, typename Sync>
__device__ __forceinline__ void fwht12_compute(T *const buf, T *const gout, const int tid64, const int lane, const unsigned long long pol_out,
                                               Sync group_sync) {
    typedef typename HPm<T>::S S;
    T v[64];
    // ---- sub-stage 1: own row, skewed chunk order
    const int k = lane & 7;
    unsigned char *const row = reinterpret_cast<unsigned char *>(buf + 64 * tid64);
    const int kb = k << 4;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const int4 q = *reinterpret_cast<const int4 *>(row + ((s << 4) ^ kb));
        v[4 * s + 0] = reinterpret_cast<const T &>(q.x); v[4 * s + 1] = reinterpret_cast<const T &>(q.y);
        v[4 * s + 2] = reinterpret_cast<const T &>(q.z); v[4 * s + 3] = reinterpret_cast<const T &>(q.w);
    }
#pragma unroll
    for (int lev = 0; lev < 2; ++lev) {                  // bits 0, 1: inside a chunk
        const int d = 1 << lev;
#pragma unroll
        for (int r = 0; r < 64; ++r)
            if ((r & d) == 0) { const T a = v[r], b = v[r | d]; v[r] = HOps<T>::add(a, b); v[r | d] = HOps<T>::sub(a, b); }
    }
#pragma unroll
    for (int lev = 0; lev < 3; ++lev) {                  // bits 2, 3, 4: chunk bits 0 .. 2, lane-dependent orientation
        const int d = 1 << lev;
        const S sg = (k & d) ? (S)-1 : (S)1;
#pragma unroll
        for (int s = 0; s < 16; ++s)
            if ((s & d) == 0) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const T a = v[4 * s + e], b = v[4 * (s | d) + e];
                    v[4 * s + e] = HPm<T>::pm(sg, a, b);
                    v[4 * (s | d) + e] = HPm<T>::pm(-sg, b, a);
                }
            }
    }
#pragma unroll
    for (int r = 0; r < 32; ++r) {                       // bit 5: chunk bit 3 (not skewed)
        const T a = v[r], b = v[r + 32];
        v[r] = HOps<T>::add(a, b); v[r + 32] = HOps<T>::sub(a, b);
    }
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        int4 q;
        q.x = reinterpret_cast<const int &>(v[4 * s + 0]); q.y = reinterpret_cast<const int &>(v[4 * s + 1]);
        q.z = reinterpret_cast<const int &>(v[4 * s + 2]); q.w = reinterpret_cast<const int &>(v[4 * s + 3]);
        *reinterpret_cast<int4 *>(row + ((s << 4) ^ kb)) = q;
    }
    group_sync();
    // ---- sub-stage 2: elements tid + 64 j
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = buf[64 * j + tid64];
#pragma unroll
    for (int lev = 0; lev < 6; ++lev) {
        const int d = 1 << lev;
#pragma unroll
        for (int r = 0; r < 64; ++r)
            if ((r & d) == 0) { const T a = v[r], b = v[r | d]; v[r] = HOps<T>::add(a, b); v[r | d] = HOps<T>::sub(a, b); }
    }
    if (pol_out) {
#pragma unroll
        for (int j = 0; j < 64; ++j) fwht_st_hint(gout + 64 * j + tid64, v[j], pol_out);
    } else {
#pragma unroll
        for (int j = 0; j < 64; ++j) gout[64 * j + tid64] = v[j];
    }
}

template <typename T> __global__ void __launch_bounds__(64, 10) fwht_first12_kernel(const __grid_constant__ FwhtFastPass p) {
    static_assert(sizeof(T) == 4, "4-byte element types only");
    extern __shared__ __align__(128) unsigned char fwht12_smem[];
    T *const buf = reinterpret_cast<T *>(fwht12_smem);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const long long tile = blockIdx.x;
    const long long col = tile >> p.log_tpc, tin = tile & (p.tiles_per_col - 1);
    const T *gin = (const T *)p.in + col * p.in_cs + (tin << 12);
    T *gout = (T *)p.out + col * p.out_cs + (tin << 12);
    const unsigned bar = fwht_smem_u32(fwht12_smem + 16384) + 8u * (unsigned)w;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fwht12_issue<T>(gin + 2048 * w, fwht_smem_u32(buf + 2048 * w), bar, p.pol_in);
    }
    __syncwarp();
    fwht_mbar_wait(bar, 0u);
    fwht12_compute<T>(buf, gout, tid, lane, p.pol_out, []() { __syncthreads(); });
}

template <typename T> struct FwhtHas12 { static constexpr bool value = false; };
template <> struct FwhtHas12<float> { static constexpr bool value = true; };
template <> struct FwhtHas12<int32_t> { static constexpr bool value = true; };

// launches the kernel above when it applies (4-byte type, 12-bit contiguous first pass, 16-byte aligned columns)
template <typename T> static int fwht_first12_launch(const FwhtFastPass &p, long long lines, cudaStream_t st, bool &done) {
    done = false;
    if constexpr (FwhtHas12<T>::value) {
        static const long off = getenv("FMB_FWHT_NO12") ? atol(getenv("FMB_FWHT_NO12")) : 0;
        if (off || p.b != 12 || p.s != 0 || p.logT != 0) return FMB_OK;
        if ((reinterpret_cast<unsigned long long>(p.in) & 15ull) || (p.in_cs & 3)) return FMB_OK;
        if (lines > 2147483647LL) return FMB_OK;
        const size_t smem = 16384 + 16;
        static int attr_done = 0;
        if (!attr_done) {
            FMB_CUDA_OK(cudaFuncSetAttribute(fwht_first12_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_done = 1;
        }
        fwht_first12_kernel<T><<<(unsigned)lines, 64, smem, st>>>(p);
        FMB_LAUNCH_OK();
        done = true;
    }
    return FMB_OK;
}


// ------------------------------------------------------------------------------------------- strided 8-bit pass, 4-byte types
// Bits [s, s + 8) of a column-major column, s >= 6: a tile is all 256 "mid" values x 64 consecutive low-index lines, and a
// thread owns FOUR neighbouring lines, so every global and shared access is 128 bits wide: a quarter of the load/store
// instructions of the one-line-per-thread kernel for the same bytes (the pass is bound by the load/store pipe, not by HBM,
// once its input comes from L2).  Shared memory is laid out [mid][line]: the sixteen threads of a mid value cover its 256
// bytes, conflict free without padding.  Two register sub-stages of four bits (ascending, as everywhere) and one exchange.
template <typename T> struct alignas(16) FwhtVec4 { T x, y, z, w; };

template <typename T> __device__ __forceinline__ void fwht16x4(FwhtVec4<T> (&v)[16]) {
#pragma unroll
    for (int lev = 0; lev < 4; ++lev) {
        const int d = 1 << lev;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            if ((r & d) == 0) {
                const FwhtVec4<T> a = v[r], b = v[r | d];
                v[r].x = HOps<T>::add(a.x, b.x); v[r].y = HOps<T>::add(a.y, b.y); v[r].z = HOps<T>::add(a.z, b.z); v[r].w = HOps<T>::add(a.w, b.w);
                v[r | d].x = HOps<T>::sub(a.x, b.x); v[r | d].y = HOps<T>::sub(a.y, b.y); v[r | d].z = HOps<T>::sub(a.z, b.z); v[r | d].w = HOps<T>::sub(a.w, b.w);
            }
        }
    }
}

template <typename T> __device__ __forceinline__ FwhtVec4<T> fwht_ld128(const FwhtVec4<T> *p) {
    int4 q;
    asm volatile("ld.global.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p));
    FwhtVec4<T> r;
    r.x = reinterpret_cast<const T &>(q.x); r.y = reinterpret_cast<const T &>(q.y); r.z = reinterpret_cast<const T &>(q.z); r.w = reinterpret_cast<const T &>(q.w);
    return r;
}

template <typename T> __device__ __forceinline__ FwhtVec4<T> fwht_ld128_cg(const FwhtVec4<T> *p) {      // L2 only: written by other SMs in this launch
    int4 q;
    asm volatile("ld.global.cg.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p) : "memory");
    FwhtVec4<T> r;
    r.x = reinterpret_cast<const T &>(q.x); r.y = reinterpret_cast<const T &>(q.y); r.z = reinterpret_cast<const T &>(q.z); r.w = reinterpret_cast<const T &>(q.w);
    return r;
}

// one tile (256 mids x 64 lines) by 256 threads; gin / gout point at (mid 0, line 4 * l16) of the tile
template <typename T, bool CG, int LV = 16>
__device__ __forceinline__ void fwht_s8_tile(const T *gin, T *gout, FwhtVec4<T> *const sm, const int tid, const long long step) {
    typedef FwhtVec4<T> V;
    const int l16 = tid & (LV - 1), q = tid / LV;
    V v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const V *src = reinterpret_cast<const V *>(gin + (long long)(16 * q + r) * step);
        v[r] = CG ? fwht_ld128_cg(src) : fwht_ld128(src);
    }
    fwht16x4<T>(v);
#pragma unroll
    for (int r = 0; r < 16; ++r) sm[(16 * q + r) * LV + l16] = v[r];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = sm[(q + 16 * r) * LV + l16];
    fwht16x4<T>(v);
#pragma unroll
    for (int r = 0; r < 16; ++r) *reinterpret_cast<V *>(gout + (long long)(q + 16 * r) * step) = v[r];
}

template <typename T, int LOGSTEP> __global__ void __launch_bounds__(256, 3) fwht_strided8_kernel(const __grid_constant__ FwhtFastPass p) {
    static_assert(sizeof(T) == 4, "4-byte element types only");
    typedef FwhtVec4<T> V;
    extern __shared__ __align__(16) unsigned char fwht_s8_smem[];
    V *const sm = reinterpret_cast<V *>(fwht_s8_smem);            // [256 mids][16 vectors]
    const int tid = threadIdx.x, l16 = tid & 15;
    const int s = LOGSTEP >= 0 ? LOGSTEP : p.s;
    const long long tile = blockIdx.x;
    const long long col = tile >> p.log_tpc, tin = tile & (p.tiles_per_col - 1);
    const int log_lo = s - 6;                                     // tiles along the low index: 2^s / 64
    const long long hi = tin >> log_lo, lo0 = (tin & (((long long)1 << log_lo) - 1)) << 6;
    const long long base = (hi << (s + 8)) + lo0 + 4 * l16;
    fwht_s8_tile<T, false>((const T *)p.in + col * p.in_cs + base, (T *)p.out + col * p.out_cs + base, sm, tid, (long long)1 << s);
}

// ------------------------------------------------------------------------------------------- order 20, persistent (experiment)
// Both passes of the order-20 transform of 4-byte types in ONE cooperative launch (FMB_FWHT_PERSIST=1): CTA c walks the
// item list c, c + G, ...; step t of the list interleaves the first-pass items of column slab t (4 lines = 64 KB each, one
// 64-thread group per line) with the second-pass tiles of slab t - D (64 KB each), so every CTA alternates between the
// passes and a waiting second pass throttles the first (tools/ubench_pipe_persistent.cu: without that back-pressure the
// first pass runs away and the intermediate is gone from L2).  A second-pass tile waits on a per-slab counter that every
// finished first-pass item bumps (release / acquire at gpu scope).  All CTAs are co-resident (cooperative launch) and take
// their items in increasing order, so the wait cannot deadlock.
struct FwhtPersistArgs {
    const void *x; void *y;
    long long xcs, ycs;
    unsigned *done;                 // [nslabs], zeroed before the launch
    int ncols, slab_cols, nslabs, dist;
    unsigned per_step;              // 2 * 64 * slab_cols
    unsigned total_items;
};

__device__ __forceinline__ unsigned fwht_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <typename T, int NT> __global__ void __launch_bounds__(NT, NT == 256 ? 2 : 5) fwht_persist20_kernel(const __grid_constant__ FwhtPersistArgs a) {
    static_assert(sizeof(T) == 4, "4-byte element types only");
    extern __shared__ __align__(128) unsigned char fwhtp_smem[];
    constexpr int NG = NT / 64, LV = NT / 16;                                      // lines per first-pass item; vectors per mid of a strided tile
    constexpr int LOG_I1 = NT == 256 ? 6 : 7, LOG_I2 = NT == 256 ? 6 : 7;          // items per column and pass: 256 / NG, 1024 / (4 LV)
    T *const buf = reinterpret_cast<T *>(fwhtp_smem);                              // NG lines / one strided tile (NT * 256 bytes)
    const unsigned bars = fwht_smem_u32(fwhtp_smem + NT * 256);                    // one mbarrier per warp
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, grp = tid >> 6, tid64 = tid & 63, w = warp & 1;
    const unsigned bar = bars + 8u * (unsigned)warp;
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1u) : "memory");
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned n1 = 0;                                                               // first-pass items done by this CTA (barrier phase)
    for (unsigned item = blockIdx.x; item < a.total_items; item += gridDim.x) {
        const unsigned step = item / a.per_step, r = item - step * a.per_step, pass2 = r & 1u, idx = r >> 1;
        const int slab = pass2 ? (int)step - a.dist : (int)step;
        if (slab < 0 || slab >= a.nslabs) continue;
        const int col = slab * a.slab_cols + (int)(idx >> LOG_I1);
        if (col >= a.ncols) continue;
        __syncthreads();                                                           // everybody is done with the buffer of the previous item
        if (!pass2) {
            const long long line = NG * (long long)(idx & ((1u << LOG_I1) - 1u)) + grp;   // line of 4096 elements inside the column
            const T *gin = (const T *)a.x + (long long)col * a.xcs + (line << 12);
            T *gout = (T *)a.y + (long long)col * a.ycs + (line << 12);
            T *const lbuf = buf + 4096 * grp;
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic accesses of the last item before the bulk copy
                fwht12_issue<T>(gin + 2048 * w, fwht_smem_u32(lbuf + 2048 * w), bar, 0ull);
            }
            __syncwarp();
            fwht_mbar_wait(bar, n1 & 1u);
            ++n1;
            fwht12_compute<T>(lbuf, gout, tid64, lane, 0ull, [grp]() { asm volatile("bar.sync %0, 64;" ::"r"(grp + 1) : "memory"); });
            __syncthreads();                                                       // all stores of the item issued
            if (tid == 0) {
                __threadfence();
                asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(a.done + slab), "r"(1u) : "memory");
            }
        } else {
            const int cols_here = min(a.slab_cols, a.ncols - slab * a.slab_cols);
            const unsigned need = (unsigned)cols_here << LOG_I1;
            if (tid == 0) while (fwht_ld_acquire(a.done + slab) < need) __nanosleep(100);
            __syncthreads();
            const unsigned tin = idx & ((1u << LOG_I2) - 1u);                      // tile in the column (order 20: one hi value): 4 LV lines each
            const long long base = (long long)tin * (4 * LV) + 4 * (tid & (LV - 1));
            T *g = (T *)a.y + (long long)col * a.ycs + base;
            fwht_s8_tile<T, true, LV>(g, g, reinterpret_cast<FwhtVec4<T> *>(buf), tid, (long long)1 << 12);
        }
    }
}

template <typename T> static int fwht_strided8_launch(const FwhtFastPass &p, long long ncols, cudaStream_t st, bool &done) {
    done = false;
    if constexpr (FwhtHas12<T>::value) {
        static const long off = getenv("FMB_FWHT_NO_S8") ? atol(getenv("FMB_FWHT_NO_S8")) : 0;
        if (off || p.b != 8 || p.s < 6 || p.order - p.s < 8) return FMB_OK;
        if ((reinterpret_cast<unsigned long long>(p.in) & 15ull) || (reinterpret_cast<unsigned long long>(p.out) & 15ull) ||
            (p.in_cs & 3) || (p.out_cs & 3)) return FMB_OK;
        FwhtFastPass q = p;
        q.log_tpc = p.order - 8 - 6;                              // tiles per column: 2^(order - 8) lines / 64
        q.tiles_per_col = (long long)1 << q.log_tpc;
        const long long grid = q.tiles_per_col * ncols;
        if (grid > 2147483647LL) return FMB_OK;
        const size_t smem = 256 * 64 * sizeof(T);
        static int attr_done = 0;
        if (!attr_done) {
            FMB_CUDA_OK(cudaFuncSetAttribute(fwht_strided8_kernel<T, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            FMB_CUDA_OK(cudaFuncSetAttribute(fwht_strided8_kernel<T, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_done = 1;
        }
        if (p.s == 12) fwht_strided8_kernel<T, 12><<<(unsigned)grid, 256, smem, st>>>(q);
        else fwht_strided8_kernel<T, -1><<<(unsigned)grid, 256, smem, st>>>(q);
        FMB_LAUNCH_OK();
        done = true;
    }
    return FMB_OK;
}

template <typename T> static int fwht_fast_launch(const FwhtFastPass &p, unsigned grid, int threads, size_t smem, cudaStream_t st) {
    if (p.b == 8 && p.s == 12) {                        // second pass of order 20 (the BASELINE shape): compile-time stride
        static int attr_done = 0;
        if (!attr_done) {
            FMB_CUDA_OK(cudaFuncSetAttribute(fwht_fast_kernel<T, 8, false, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 << 10));
            attr_done = 1;
        }
        fwht_fast_kernel<T, 8, false, 12><<<grid, threads, smem, st>>>(p);
        FMB_LAUNCH_OK();
        return FMB_OK;
    }
#define FMB_FWHT_CASE(BB)                                                                                              \
    case BB: {                                                                                                         \
        static int attr_done = 0;                                                                                      \
        if (!attr_done) {                                                                                              \
            FMB_CUDA_OK(cudaFuncSetAttribute(fwht_fast_kernel<T, BB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 << 10)); \
            FMB_CUDA_OK(cudaFuncSetAttribute(fwht_fast_kernel<T, BB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 << 10)); \
            attr_done = 1;                                                                                             \
        }                                                                                                              \
        if (p.s == 0) fwht_fast_kernel<T, BB, true><<<grid, threads, smem, st>>>(p);                                   \
        else fwht_fast_kernel<T, BB, false><<<grid, threads, smem, st>>>(p);                                           \
        break;                                                                                                         \
    }
    switch (p.b) {
        FMB_FWHT_CASE(4) FMB_FWHT_CASE(5) FMB_FWHT_CASE(6) FMB_FWHT_CASE(7) FMB_FWHT_CASE(8)
        FMB_FWHT_CASE(9) FMB_FWHT_CASE(10) FMB_FWHT_CASE(11) FMB_FWHT_CASE(12)
        default: set_error("FWHT fast path: unsupported pass width %d", p.b); return FMB_ERR_VALUE;
    }
#undef FMB_FWHT_CASE
    FMB_LAUNCH_OK();
    return FMB_OK;
}


template <typename T>
static int fwht_persist20_launch(const void *x, int64_t xcs, void *y, int64_t ycs, int64_t M, void *ws, int64_t ws_bytes, cudaStream_t st, bool &done) {
    done = false;
    if constexpr (FwhtHas12<T>::value) {
        static const long slab_env = getenv("FMB_FWHT_PERSIST_SLAB") ? atol(getenv("FMB_FWHT_PERSIST_SLAB")) : 2;
        static const long dist_env = getenv("FMB_FWHT_PERSIST_DIST") ? atol(getenv("FMB_FWHT_PERSIST_DIST")) : 2;
        static const long ctas_env = getenv("FMB_FWHT_PERSIST_CTAS") ? atol(getenv("FMB_FWHT_PERSIST_CTAS")) : 8;
        if ((reinterpret_cast<unsigned long long>(x) & 15ull) || (reinterpret_cast<unsigned long long>(y) & 15ull) || (xcs & 3) || (ycs & 3)) return FMB_OK;
        if (x == (const void *)y || M >= ((int64_t)1 << 24)) return FMB_OK;
        FwhtPersistArgs a;
        memset(&a, 0, sizeof(a));
        a.x = x; a.y = y; a.xcs = xcs; a.ycs = ycs;
        a.ncols = (int)M; a.slab_cols = (int)std::max<long>(1, std::min<long>(slab_env, 16));
        a.nslabs = (int)((M + a.slab_cols - 1) / a.slab_cols);
        a.dist = (int)std::max<long>(1, dist_env);
        static const long nt_env = getenv("FMB_FWHT_PERSIST_NT") ? atol(getenv("FMB_FWHT_PERSIST_NT")) : 256;
        const int NT = nt_env == 128 ? 128 : 256;
        a.per_step = 2u * (NT == 256 ? 64u : 128u) * (unsigned)a.slab_cols;
        const int64_t total = ((int64_t)a.nslabs + a.dist) * a.per_step;
        if (total >= ((int64_t)1 << 32) || ws == nullptr || ws_bytes < (int64_t)a.nslabs * 4) return FMB_OK;
        a.total_items = (unsigned)total;
        a.done = (unsigned *)ws;
        const size_t smem = (size_t)NT * 256 + 64;
        const void *kern = NT == 256 ? (const void *)fwht_persist20_kernel<T, 256> : (const void *)fwht_persist20_kernel<T, 128>;
        static int per_sm = 0;
        if (!per_sm) {
            FMB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            FMB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
            if (per_sm < 1) per_sm = -1;
        }
        if (per_sm < 1) return FMB_OK;
        int grid = device_props().sm_count * (int)std::min<long>(per_sm, std::max<long>(1, ctas_env));
        if ((grid & 1) == 0) --grid;                                 // odd: every CTA alternates between the two passes
        if ((unsigned)grid > a.total_items) grid = (int)a.total_items | 1;
        FMB_CUDA_OK(cudaMemsetAsync(ws, 0, (size_t)a.nslabs * 4, st));
        void *args[1] = {&a};
        const cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)grid), dim3((unsigned)NT), args, smem, st);
        if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources) { (void)cudaGetLastError(); return FMB_OK; }
        if (e != cudaSuccess) { set_error("persistent FWHT launch failed: %s", cudaGetErrorString(e)); return FMB_ERR_CUDA; }
        g_launches.fetch_add(1);
        done = true;
    }
    return FMB_OK;
}

template <typename T>
static int fwht_fast(int order, const void *x, int64_t xcs, void *y, int64_t ycs, int64_t M, void *ws, int64_t ws_bytes, cudaStream_t st) {
    {
        static const long persist = getenv("FMB_FWHT_PERSIST") ? atol(getenv("FMB_FWHT_PERSIST")) : 0;
        if (persist && order == 20 && M >= 16) {
            bool done = false;
            int rc = fwht_persist20_launch<T>(x, xcs, y, ycs, M, ws, ws_bytes, st, done);
            if (rc || done) return rc;
        }
    }
    // bit ranges per pass: a contiguous first pass of up to 12 bits, then strided passes of 4..8 bits
    std::vector<int> bits;
    if (order <= 12) bits.push_back(order);
    else {
        int rem = order - 12;
        if (rem < 4) { bits.push_back(order - 4); bits.push_back(4); }
        else {
            bits.push_back(12);
            int np = (rem + 7) / 8, basev = rem / np, extra = rem % np;
            for (int i = 0; i < np; ++i) bits.push_back(basev + (i < extra ? 1 : 0));
        }
    }
    const int tile_elems = (int)std::min<size_t>(8192, std::max<size_t>(512, (size_t)32768 / sizeof(T)));   // <= 512 threads x 16
    size_t l2 = device_props().l2_bytes ? device_props().l2_bytes : (size_t)100 << 20;
    static const long slab_mb = getenv("FMB_FWHT_SLAB_MB") ? atol(getenv("FMB_FWHT_SLAB_MB")) : 0;
    static const long pipe_ns = getenv("FMB_FWHT_PIPE_STREAMS") ? atol(getenv("FMB_FWHT_PIPE_STREAMS")) : 3;
    static const long pipe_mb = getenv("FMB_FWHT_PIPE_MB") ? atol(getenv("FMB_FWHT_PIPE_MB")) : 16;
    // Multi-pass orders: "pipelined slabs" (common.h: PipeScope) - slabs of a few columns whose intermediate (y itself:
    // the passes after the first work in place) is still in L2 when the next pass reads it, issued round-robin on
    // internal streams so that launch gaps and partial waves of one slab are filled by the next.  Fallback (few columns,
    // or FMB_FWHT_PIPE_STREAMS=1): one launch per pass over a 512 MiB slab.
    (void)l2;
    const size_t col_bytes = ((size_t)1 << order) * sizeof(T);
    int ns = 1;
    int64_t slab;
    {
        size_t budget = slab_mb > 0 ? (size_t)slab_mb << 20 : (size_t)512 << 20;
        slab = (int64_t)(budget / col_bytes);
        if (slab < 1) slab = 1;
        if (bits.size() > 1 && pipe_ns > 1 && slab_mb == 0) {
            int64_t c = std::max<int64_t>(1, (int64_t)(((size_t)pipe_mb << 20) / col_bytes));
            if (M >= 2 * c * pipe_ns) { slab = c; ns = (int)pipe_ns; }
        }
    }
    if (bits.size() == 1) slab = M;
    PipeScope pipe;
    {
        int rc = pipe.begin(ns, st);
        if (rc) return rc;
    }
    int64_t slab_idx = 0;
    for (int64_t c0 = 0; c0 < M; c0 += slab, ++slab_idx) {
        const int64_t nc = std::min<int64_t>(slab, M - c0);
        st = pipe.stream(slab_idx);
        int s = 0;
        for (size_t pi = 0; pi < bits.size(); ++pi) {
            FwhtFastPass p;
            memset(&p, 0, sizeof(p));
            const bool first = (pi == 0);
            p.in = first ? (const char *)x + (size_t)(c0 * xcs) * sizeof(T) : (const char *)y + (size_t)(c0 * ycs) * sizeof(T);
            p.in_cs = first ? xcs : ycs;
            p.out = (char *)y + (size_t)(c0 * ycs) * sizeof(T);
            p.out_cs = ycs;
            p.b = bits[pi]; p.s = s; p.order = order;
            // L2 eviction policies of the multi-pass pipeline (FMB_FWHT_HINT, bit mask): 1 = the column input is read
            // evict-first, 2 = intermediates are written evict-last, 4 = intermediates are read evict-first (they are
            // overwritten in place), 8 = the final result is written evict-first
            {
                static const long hint = getenv("FMB_FWHT_HINT") ? atol(getenv("FMB_FWHT_HINT")) : 0;
                const unsigned long long EV_FIRST = 0x12F0000000000000ull, EV_LAST = 0x14F0000000000000ull;
                const bool last = (pi + 1 == bits.size());
                if (bits.size() > 1) {
                    if (first && (hint & 1)) p.pol_in = EV_FIRST;
                    if (!first && (hint & 4)) p.pol_in = EV_FIRST;
                    if (!last && (hint & 2)) p.pol_out = EV_LAST;
                    if (last && (hint & 8)) p.pol_out = EV_FIRST;
                }
            }
            int logT = 0;
            if (s > 0) {
                int want = tile_elems >> p.b;                      // lines per tile
                while ((1 << (logT + 1)) <= want && (logT + 1) <= s) ++logT;
            }
            p.logT = logT;
            const int threads = (1 << (p.b - 4)) << logT;
            p.tiles_per_col = ((long long)1 << (order - p.b)) >> logT;
            p.log_tpc = order - p.b - logT;
            const long long grid = p.tiles_per_col * nc;
            if (grid > 2147483647LL || threads > 512 || threads < 1) { set_error("FWHT fast path: bad geometry"); return FMB_ERR_VALUE; }
            const size_t smem = (size_t)(1 << logT) * ((size_t)(1 << p.b) + ((size_t)(1 << p.b) >> 4) + 1) * sizeof(T);
            bool done12 = false;
            int rc = fwht_first12_launch<T>(p, grid, st, done12);
            if (rc) return rc;
            if (!done12 && (rc = fwht_strided8_launch<T>(p, nc, st, done12))) return rc;
            if (!done12 && (rc = fwht_fast_launch<T>(p, (unsigned)grid, threads, smem, st))) return rc;
            s += p.b;
        }
    }
    return pipe.end();
}

template <typename T>
static int fwht_typed(int order, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs, int64_t ycs, int64_t M, void *ws,
                      int64_t ws_bytes, cudaStream_t st) {
    const bool row_major = (xcs == 1 && M > 1);
#ifndef FMB_EMULATE
    static const long no_fast = getenv("FMB_NO_FAST") ? atol(getenv("FMB_NO_FAST")) : 0;
    if (!no_fast && !row_major && xrs == 1 && yrs == 1 && order >= 8 && order <= 40)
        return fwht_fast<T>(order, x, xcs, y, ycs, M, ws, ws_bytes, st);
#endif
    // ---- split the bits into passes
    const int tile_bytes = 64 << 10;
    std::vector<int> bits;
    int remaining = order;
    int max_strided_T = (sizeof(T) >= 8) ? 16 : 32;
    int bmax_strided = 0;
    while (((size_t)1 << (bmax_strided + 1)) * max_strided_T * sizeof(T) <= (size_t)tile_bytes && bmax_strided < 10) ++bmax_strided;
    if (!row_major) {
        int b0max = 0;
        while (((size_t)1 << (b0max + 1)) * sizeof(T) <= (size_t)(tile_bytes / 2) && b0max < 12) ++b0max;
        int b0 = std::min(order, b0max);
        bits.push_back(b0);
        remaining -= b0;
    }
    if (remaining > 0) {
        int npass = (remaining + bmax_strided - 1) / bmax_strided;
        int base = remaining / npass, extra = remaining % npass;
        for (int i = 0; i < npass; ++i) bits.push_back(base + (i < extra ? 1 : 0));
    }
    // ---- slab of columns that stays L2-resident between passes
    size_t l2 = device_props().l2_bytes ? device_props().l2_bytes : (size_t)100 << 20;
    int64_t slab = (int64_t)((l2 / 3) / (((size_t)1 << order) * sizeof(T)));
    if (slab < 1) slab = 1;
    if (bits.size() == 1) slab = M;
    for (int64_t c0 = 0; c0 < M; c0 += slab) {
        const int64_t nc = std::min<int64_t>(slab, M - c0);
        int s = 0;
        for (size_t pi = 0; pi < bits.size(); ++pi) {
            FwhtPass p;
            memset(&p, 0, sizeof(p));
            p.b = bits[pi]; p.s = s;
            p.row_major = row_major;
            p.ncols = nc;
            p.hi_count = (long long)1 << (order - s - p.b);
            if (row_major) {
                p.lo_count = (long long)1 << s;
                p.mid_contig = 0;
                p.T = max_strided_T;
                p.lines_total = nc * p.lo_count * p.hi_count;
                p.pad = 0;
            } else if (s == 0) {
                p.lo_count = 1;
                p.mid_contig = 1;
                p.T = 1;
                // small transforms: several columns / chunks per CTA so that a CTA has >= 2048 elements of work
                while (((long long)p.T << p.b) < 2048 && p.T < 256) p.T *= 2;
                p.lines_total = nc * p.hi_count;
                p.pad = 1;
            } else {
                p.lo_count = (long long)1 << s;
                p.mid_contig = 0;
                p.T = (int)std::min<long long>(max_strided_T, p.lo_count);
                p.lines_total = nc * p.lo_count * p.hi_count;
                p.pad = (p.T < 32) ? 1 : 0;
            }
            const bool first = (pi == 0);
            p.in = first ? (const char *)x + (size_t)(c0 * xcs) * sizeof(T) : (const char *)y + (size_t)(c0 * ycs) * sizeof(T);
            p.in_rs = first ? xrs : yrs; p.in_cs = first ? xcs : ycs;
            p.out = (char *)y + (size_t)(c0 * ycs) * sizeof(T);
            p.out_rs = yrs; p.out_cs = ycs;
            int rc = launch_fwht_pass<T>(p, st);
            if (rc) return rc;
            s += p.b;
        }
    }
    return FMB_OK;
}

int fwht_apply(int order, const void *x, int64_t xrs, int64_t xcs, void *y, int64_t yrs, int64_t ycs, int64_t M, int dtype, void *ws,
               int64_t ws_bytes, cudaStream_t st) {
    switch (dtype) {
        case FMB_INT8: return fwht_typed<int8_t>(order, x, xrs, xcs, y, yrs, ycs, M, ws, ws_bytes, st);
        case FMB_INT16: return fwht_typed<int16_t>(order, x, xrs, xcs, y, yrs, ycs, M, ws, ws_bytes, st);
        case FMB_INT32: return fwht_typed<int32_t>(order, x, xrs, xcs, y, yrs, ycs, M, ws, ws_bytes, st);
        case FMB_INT64: return fwht_typed<int64_t>(order, x, xrs, xcs, y, yrs, ycs, M, ws, ws_bytes, st);
        case FMB_FLOAT32: return fwht_typed<float>(order, x, xrs, xcs, y, yrs, ycs, M, ws, ws_bytes, st);
        case FMB_FLOAT64: return fwht_typed<double>(order, x, xrs, xcs, y, yrs, ycs, M, ws, ws_bytes, st);
        case FMB_COMPLEX64: return fwht_typed<float2>(order, x, xrs, xcs, y, yrs, ycs, M, ws, ws_bytes, st);
        case FMB_COMPLEX128: return fwht_typed<double2>(order, x, xrs, xcs, y, yrs, ycs, M, ws, ws_bytes, st);
        default: set_error("Hadamard: unsupported dtype %d", dtype); return FMB_ERR_TYPE;
    }
}

}  // namespace fmb
