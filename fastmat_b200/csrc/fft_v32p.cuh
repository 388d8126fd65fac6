// "V32P": ALL passes of a 2^20-point complex64 transform / convolution in ONE persistent launch, input tiles brought in
// by the TMA engine ahead of the arithmetic.
//
// The per-pass kernels of fft_v32.cuh spend most of their time waiting: a CTA loads its tile (32 LDG per thread), then
// computes, then stores, and with 128 registers per thread only two CTAs fit on an SM, so the load latency of one tile
// is hidden by at most one other tile.  Here one CTA of 18 working warps stays resident per SM
// (launched as 20: register files are allocated per four warps, which leaves 96 registers per thread - enough):
//
//   * warps 0-7 and 8-15 are two compute "groups" (256 threads = 8 lines x 32 threads, the V32 tile shape) working on
//     different tiles.  The tile data does not pass through registers on the way in: it is waited for on an mbarrier
//     and read from one of THREE 66 KB shared-memory buffers, which is then reused in place as the exchange buffer
//     between the two radix-32 stages.  Compute warps never wait for anything but their data (mbarrier.try_wait; two
//     barriers per buffer, see v32p_kernel).
//   * warp 16 (one lane) is the requester: as soon as all eight warps of a group have handed a buffer back (per-warp
//     counters in shared memory) it waits for the next tile's producers (global counters, see below) and issues the
//     bulk-tensor copies - cp.async.bulk.tensor, 4 boxes of 256 rows x 64 B, for the strided passes (zero padding =
//     out-of-bounds rows of the tensor map); cp.async.bulk, 8 lines of 8 KB, for the contiguous pass.  A request is in
//     flight during the whole arithmetic of the two tiles being worked on.
//   * warp 17 (one lane) is the signaller: when all eight warps of a group have issued the stores of a tile it
//     publishes the tile's completion at GPU scope (release reduction: the fence is paid by a warp that has nothing
//     else to do).
//
// Work list (static assignment): item w is tile t of pass p of column slab s, ordered
// A(s), M(s-D), C(s-2D) per step so that a slab's intermediate is produced and consumed while it is still in the L2;
// CTA c owns items c, c+G, c+2G, ...  A tile of pass p may be requested once all tiles of pass p-1 of its slab have
// signalled, a tile of pass A once the last pass of slab s - NSLOT has released the ring slot.  The requester and the
// signaller never block each other and each CTA works through its items in order, so the unfinished item with the
// smallest index can always make progress: the polls cannot deadlock as long as the G CTAs are co-resident (one per
// SM).  The delay D and the ring size are chosen on the host such that producers are normally several items per CTA
// ahead of their consumers (fft_engine.cu: v32p_geom).
#pragma once
#include <cuda.h>

#include "fft_v32.cuh"

namespace fmb {

constexpr int V32P_NT = 640, V32P_NBUF = 3;                // 16 compute warps + one warpgroup for requester and signaller
constexpr unsigned V32P_TILE_BYTES = 65536;                                     // 8 lines x 1024 complex64
constexpr size_t V32P_BUF = V32_SMEM;                                           // 67712 = 529 * 128
constexpr size_t V32P_CTRL = 256;                                               // mbarriers and counters, see v32p_kernel
constexpr size_t V32P_SMEM = V32P_NBUF * V32P_BUF + 128 /* alignment slack */ + V32P_CTRL;
static_assert(V32P_BUF % 128 == 0, "buffers must stay 128-byte aligned for the tensor copies");

struct V32PArgs {
    FastArgs<float2> pass[3];      // .out of the last pass = y (slab 0); everything about the ring is below
    int npass;                     // 2: plain transform, 3: convolution
    int ncols, slab_cols, nslabs, delay, nslot, mix;
    unsigned tiles;                // tiles per full slab and pass = slab_cols * 128
    unsigned items_per_step, total_items;
    long long slot_stride;         // elements between ring slots = slab_cols * L
    long long ring_cs;             // elements between the columns of a slot (L; in-place variants: the ring IS y, slot = slab)
    long long y_slab_stride;       // elements between the outputs of consecutive slabs
    long long L;
    float2 *ring;
    unsigned *done;                // [npass][nslabs] completed tiles
    unsigned long long hint_x, hint_ring;      // L2 cache policies of the two kinds of tile request
};

// Flag protocol WITHOUT L1 invalidation.  A gpu-scope acquire load or __threadfence() makes ptxas emit CCTL.IVALL, which
// throws away the SM's whole L1 (twiddle tables included) once per tile.  It is not needed here: the data guarded by the
// flags (the ring slots) is only ever read through the async proxy / from L2, never from L1.  So the consumer polls with
// a relaxed load and the producer publishes with a release reduction (MEMBAR.ALL.GPU + RED, no CCTL).
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned v32p_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void v32p_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void v32p_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// blocking wait; the time hint lets the hardware park the warp instead of returning at once (a bare try_wait loop
// retried ~200 times per tile and took a quarter of all issued instructions in the convolution kernel)
__device__ __forceinline__ void v32p_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "V32P_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra V32P_DONE_%=;\n\t"
        "bra V32P_WAIT_%=;\n\t"
        "V32P_DONE_%=:\n\t}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
// 3-D box {8 elements, 256 rows, 1 column} of a tensor of 8-byte elements -> 16 KB of shared memory
__device__ __forceinline__ void v32p_tma_box(unsigned dst, const CUtensorMap *map, unsigned bar, int c0, int c1, int c2,
                                             unsigned long long hint) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(hint) : "memory");
}
__device__ __forceinline__ void v32p_bulk_line(unsigned dst, const void *src, unsigned bytes, unsigned bar, unsigned long long hint) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(hint) : "memory");
}
__device__ __forceinline__ void v32p_group_bar(int group) {
    asm volatile("bar.sync %0, 256;" ::"r"(group + 1) : "memory");
}

__device__ __forceinline__ unsigned v32p_ld_acq(unsigned addr) {
    unsigned v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void v32p_st_rel(unsigned addr, unsigned v) {
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

struct V32PItem { int p, slab; unsigned tile; bool in_range, compute; };
__device__ __forceinline__ V32PItem v32p_decode(const V32PArgs &g, unsigned w) {
    V32PItem it;
    const unsigned step = w / g.items_per_step, r = w - step * g.items_per_step;
    if (g.mix) { it.p = (int)(r % (unsigned)g.npass); it.tile = r / (unsigned)g.npass; }
    else { it.p = (int)(r / g.tiles); it.tile = r - (unsigned)it.p * g.tiles; }
    it.slab = (int)step - it.p * g.delay;
    it.in_range = it.slab >= 0 && it.slab < g.nslabs;
    const int cols_here = min(g.slab_cols, g.ncols - it.slab * g.slab_cols);
    it.compute = it.in_range && (int)(it.tile >> 7) < cols_here;       // ragged last slab: tile is signalled, not computed
    return it;
}

// Has everything item w reads been produced (and, for pass A, has the ring slot it writes been released)?  One poll.
template <bool CONV> __device__ __forceinline__ bool v32p_ready(const V32PArgs &g, const V32PItem &it) {
    constexpr int LAST = CONV ? 2 : 1;
    if (!it.compute) return true;
    const unsigned *flag = nullptr;
    if (it.p == 0) {
        const int prev = it.slab - g.nslot;
        if (prev >= 0) flag = g.done + (size_t)LAST * g.nslabs + prev;
    } else flag = g.done + (size_t)(it.p - 1) * g.nslabs + it.slab;
    return flag == nullptr || ld_relaxed_gpu(flag) >= g.tiles;
}

// Executed by ONE thread: request the tile of item `it` into `buf` (completion on `bar`).  Items that are not computed
// (slabs outside the batch at both ends of the list, columns beyond a ragged last slab) still move 64 KB from a valid
// address, which keeps the buffer / barrier protocol free of special cases.
template <bool CONV, bool BOX1>
__device__ __forceinline__ void v32p_request(const V32PArgs &g, const CUtensorMap *map_x, const CUtensorMap *map_ring, const V32PItem &it,
                                             unsigned buf, unsigned bar) {
    constexpr int LAST = CONV ? 2 : 1;
    const int slab = it.in_range ? it.slab : 0;
    const unsigned col = it.compute ? (it.tile >> 7) : 0u, i0 = (it.tile & 127u) << V32_LOGT;
    const int slot = slab % g.nslot;
    asm volatile("fence.proxy.async;" ::: "memory");      // the data behind the flag is read through the async proxy
    v32p_mbar_expect_tx(bar, V32P_TILE_BYTES);
    if (it.p == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            v32p_tma_box(buf + q * 16384u, map_x, bar, (int)i0, 256 * q, slab * g.slab_cols + (int)col, g.hint_x);
    } else if ((CONV && it.p == LAST) || (!CONV && BOX1)) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            v32p_tma_box(buf + q * 16384u, map_ring, bar, (int)i0, 256 * q, slot * g.slab_cols + (int)col, g.hint_ring);
    } else {
        const float2 *src = g.ring + (long long)slot * g.slot_stride + (long long)col * g.ring_cs + (long long)i0 * 1024;
#pragma unroll
        for (int t = 0; t < V32_T; ++t)
            v32p_bulk_line(buf + (unsigned)(t * V32_RS * sizeof(float2)), src + t * 1024, 8192u, bar, g.hint_ring);
    }
}

// One tile of one pass by one group of 256 threads; `buf` holds the tile (strided passes: [f][8 lines] dense; contiguous
// pass: line t at t * V32_RS) and becomes the exchange buffer.  `handback()` is called once the buffer is no longer read.
template <unsigned OPT, typename Handback>
__device__ __forceinline__ void v32p_tile(const FastArgs<float2> &a, float2 *const smem, float2 *const out_col, const unsigned i0,
                                          const int gt, const int group, Handback handback) {
    typedef float2 C;
    constexpr bool LOAD_T = (OPT & FO_LOAD_T) != 0, STORE_T = (OPT & FO_STORE_T) != 0, TWO = (OPT & FO_TWO_FFTS) != 0;
    static_assert(!(OPT & (FO_IN_MASK | FO_PRE | FO_POST)), "zero padding is done by the tensor copy; no Bluestein here");
    C v[32];
    int jb, t;
    v32_pos<LOAD_T>(gt, jb, t);
    {
        const C *sl = LOAD_T ? smem + (jb << V32_LOGT) + t : smem + t * V32_RS + jb;
        V32Chain h;
        if (OPT & FO_IN_TWIDDLE) h = v32_chain_init(a, i0 + t, jb);
#pragma unroll
        for (int m = 0; m < 32; ++m) {
            C val = LOAD_T ? sl[(32 * m) << V32_LOGT] : sl[32 * m];
            if (OPT & FO_IN_CONJ) val = cconj(val);
            if (OPT & FO_IN_TWIDDLE) {
                val = cmul(val, h.c[m & 3]);
                if (m + 4 < 32) h.c[m & 3] = cmul(h.c[m & 3], h.s4);
            }
            v[m] = val;
        }
    }
    dft32(v);
    // in place: everybody who reads a region must have done so before it is overwritten in the exchange layout
    if constexpr (LOAD_T) v32p_group_bar(group);
    else __syncwarp();
    {
        C *sl = smem + t * V32_RS + jb * 33;
#pragma unroll
        for (int q = 0; q < 32; ++q) sl[q] = v[q];
    }

    const CPair<C> *const tab = reinterpret_cast<const CPair<C> *>(a.tw);
    auto stage_b = [&](int jb_, int t_, const CPair<C> *tb) {
        const C *sl = smem + t_ * V32_RS + jb_;
#pragma unroll
        for (int m = 0; m < 32; ++m) v[m] = sl[33 * m];
        const CPair<C> *tp = tb + jb_;
        dft32_stage_tw<C, false>(v, tp);           // the same arithmetic as the per-pass kernels (bit-identical outputs)
    };
    auto final_store = [&](int jb_, int t_) {
        const unsigned i = i0 + t_;
        // fixed 1024 x 1024 geometry (see v32_pass_kernel): one base register, immediate offsets
        C *dst = out_col + (STORE_T ? (int)i + jb_ * 1024 : (int)i * 1024 + jb_);
        constexpr int kstep = 32 * 1024;
        V32Chain h;
        if (OPT & FO_TWIDDLE) h = v32_chain_init(a, i, jb_);
        int klim = 0;
        if (OPT & FO_OUT_MASK) {
            const int room = a.out_n - (int)i * a.out_li;
            klim = (room > 0 ? (room + a.out_lk - 1) / a.out_lk : 0) - jb_;
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            C val = v[q];
            if (OPT & FO_TWIDDLE) {
                val = cmul(val, h.c[q & 3]);
                if (q + 4 < 32) h.c[q & 3] = cmul(h.c[q & 3], h.s4);
            }
            if (OPT & FO_OUT_CONJ) val = cconj(val);
            bool ok = true;
            if (OPT & FO_OUT_MASK) ok = 32 * q < klim;
            C *ps = STORE_T ? dst + q * kstep : dst + 32 * q;
            if (ok) *ps = val;
        }
    };

    if constexpr (!TWO) {
        if constexpr (!LOAD_T && !STORE_T) __syncwarp();
        else v32p_group_bar(group);
        v32_pos<STORE_T>(gt, jb, t);
        stage_b(jb, t, tab);
        handback();
        final_store(jb, t);
    } else {
        static_assert(!LOAD_T && !STORE_T, "the middle pass works on contiguous, warp-private lines");
        __syncwarp();
        const C *mp = a.mid + (long long)(i0 + t) * a.mid_is + jb;
        C mh[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) mh[r] = ld_nc_ordered(mp + 32 * (2 * r));
        stage_b(jb, t, tab);
        {
            // the same arithmetic as v32_pass_kernel's middle pass: the spectrum rides on the first butterfly level
            auto sw = [](C m) { return (OPT & FO_MID_CONJ) ? m : cconj(m); };
            C e[16], o[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) e[r] = cconj(v[2 * r]);
#pragma unroll
            for (int r0 = 0; r0 < 4; ++r0)
                dft4_tw4(e[r0], e[4 + r0], e[8 + r0], e[12 + r0], sw(mh[r0]), sw(mh[4 + r0]), sw(mh[8 + r0]), sw(mh[12 + r0]));
#pragma unroll
            for (int r = 0; r < 16; ++r) mh[r] = ld_nc_ordered(mp + 32 * (2 * r + 1));
            __syncwarp();
            dft16_level2(e);
#pragma unroll
            for (int r = 0; r < 16; ++r) o[r] = cconj(v[2 * r + 1]);
#pragma unroll
            for (int r0 = 0; r0 < 4; ++r0)
                dft4_tw4(o[r0], o[4 + r0], o[8 + r0], o[12 + r0], sw(mh[r0]), sw(mh[4 + r0]), sw(mh[8 + r0]), sw(mh[12 + r0]));
            dft16_level2(o);
            dft32_combine<C, 0>(v, e, o);
        }
        {
            C *sl = smem + t * V32_RS + jb * 33;
#pragma unroll
            for (int q = 0; q < 32; ++q) sl[q] = v[q];
        }
        __syncwarp();
        stage_b(jb, t, tab + 512);
        handback();
        final_store(jb, t);
    }
}

template <unsigned OA, unsigned OB, unsigned OC>
__global__ void __launch_bounds__(V32P_NT, 1)
v32p_kernel(const __grid_constant__ V32PArgs g, const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_ring) {
    typedef float2 C;
    constexpr bool CONV = OC != 0;
    constexpr int LAST = CONV ? 2 : 1;
    extern __shared__ unsigned char v32p_smem_raw[];
    unsigned char *const base = v32p_smem_raw + ((128u - (v32p_smem_u32(v32p_smem_raw) & 127u)) & 127u);
    const unsigned base_s = v32p_smem_u32(base);
    // control block: full[3][2] mbarriers (TMA completion) | handed[16], stored[16]: per compute warp, how many of its
    // items have handed their buffer back / issued their stores.
    // The k-th use of buffer j completes on full[j][k & 1] (phase k >> 1).  try_wait.parity cannot tell "phase not started"
    // from "completed two phases ago", so a warp must only test a barrier whose PREVIOUS phase it has consumed itself:
    // consecutive uses of a buffer alternate between the two groups, uses k and k - 2 belong to the same group.
    const unsigned ctrl = base_s + (unsigned)(V32P_NBUF * V32P_BUF);
    const unsigned full_s = ctrl, handed_s = ctrl + 64, stored_s = ctrl + 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 64) reinterpret_cast<unsigned *>(base + V32P_NBUF * V32P_BUF)[tid] = 0u;
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int j = 0; j < 2 * V32P_NBUF; ++j) v32p_mbar_init(full_s + 8 * j, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned G = gridDim.x;
    const unsigned nloc = g.total_items > blockIdx.x ? (g.total_items - blockIdx.x + G - 1) / G : 0u;

    // (register files are allocated in units of four warps: 18 warps cost as much as 20, i.e. 96 registers per thread)
    if (warp > 17) return;
    if (warp == 16) {
        // ------------------------------------------------------------------ requester
        for (unsigned n = 0; n < nloc; ++n) {
            const V32PItem it = v32p_decode(g, blockIdx.x + n * G);
            const unsigned j = n % V32P_NBUF;
            // producers first (normally long done: one poll through L2), so that nothing but the issue itself is left to
            // do at the moment the buffer comes back
            if (lane == 0) while (!v32p_ready<CONV>(g, it)) __nanosleep(64);
            if (n >= (unsigned)V32P_NBUF) {           // buffer j was used by item n - 3: all eight warps of its group done with it?
                const unsigned prev = n - V32P_NBUF, gp = prev & 1u, kp = prev >> 1;
                for (;;) {
                    const unsigned cnt = lane < 16 ? v32p_ld_acq(handed_s + 4 * lane) : 0xffffffffu;
                    const bool ok = (lane >> 3) != gp || cnt > kp;
                    if (__all_sync(0xffffffffu, ok)) break;
                    __nanosleep(20);
                }
                __syncwarp();                         // lane 0 inherits what the other lanes acquired
            }
            if (lane == 0) {
                v32p_request<CONV, (OB & FO_LOAD_T) != 0>(g, &map_x, &map_ring, it, base_s + j * (unsigned)V32P_BUF,
                                   full_s + 8 * (2 * j + ((n / V32P_NBUF) & 1u)));
            }
            __syncwarp();
        }
        return;
    }
    if (warp == 17) {
        // ------------------------------------------------------------------ signaller
        for (unsigned n = 0; n < nloc; ++n) {
            const V32PItem it = v32p_decode(g, blockIdx.x + n * G);
            const unsigned gs = n & 1u, k = n >> 1;
            for (;;) {
                const unsigned cnt = lane < 16 ? v32p_ld_acq(stored_s + 4 * lane) : 0xffffffffu;
                const bool ok = (lane >> 3) != gs || cnt > k;
                if (__all_sync(0xffffffffu, ok)) break;
                __nanosleep(64);
            }
            __syncwarp();                             // lane 0 inherits what the other lanes acquired
            if (lane == 0 && it.in_range) red_release_gpu(g.done + (size_t)it.p * g.nslabs + it.slab, 1u);
            __syncwarp();
        }
        return;
    }

    // ---------------------------------------------------------------------- compute groups
    const int group = tid >> 8, gt = tid & 255;
    for (unsigned n = group; n < nloc; n += 2) {
        const V32PItem it = v32p_decode(g, blockIdx.x + n * G);
        const unsigned j = n % V32P_NBUF, k = n >> 1;
        C *const buf = reinterpret_cast<C *>(base + j * V32P_BUF);
        const unsigned use = n / V32P_NBUF;
        v32p_mbar_wait(full_s + 8 * (2 * j + (use & 1u)), (use >> 1) & 1u);
        auto handback = [&]() {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) v32p_st_rel(handed_s + 4 * warp, k + 1);
        };
        if (it.compute) {
            const unsigned col = it.tile >> 7, i0 = (it.tile & 127u) << V32_LOGT;
            const int slot = it.slab % g.nslot;
            C *const ring_col = g.ring + (long long)slot * g.slot_stride + (long long)col * g.ring_cs;
            if (it.p == 0) v32p_tile<OA>(g.pass[0], buf, ring_col, i0, gt, group, handback);
            else if (it.p == LAST) {
                C *const y_col = g.pass[LAST].out + (long long)it.slab * g.y_slab_stride + (long long)col * g.pass[LAST].out_cs;
                if constexpr (CONV) v32p_tile<OC>(g.pass[2], buf, y_col, i0, gt, group, handback);
                else v32p_tile<OB>(g.pass[1], buf, y_col, i0, gt, group, handback);
            } else {
                if constexpr (CONV) v32p_tile<OB>(g.pass[1], buf, ring_col, i0, gt, group, handback);
            }
        } else handback();
        __syncwarp();                                   // every lane has issued its stores
        if (lane == 0) v32p_st_rel(stored_s + 4 * warp, k + 1);
    }
}

// ---- "V32T": ONE strided pass as an ordinary (non-persistent) launch whose tile is fetched by the TMA engine: thread 0
// requests the four 16 KB boxes of the CTA's tile, everybody waits on the mbarrier and runs the same tile code as the
// persistent kernel (landing buffer = exchange buffer).  Against the register-direct loads of v32_pass_kernel this takes
// the 64-byte-segment loads (four L1 wavefronts per warp instruction) off the load/store pipe and leaves the registers
// free while the tile is in flight; zero padding (Toeplitz) is the out-of-bounds fill of the tensor map.  Used by the
// pipelined-slab schedule for the first and the last pass of a convolution (fft_engine.cu: run_v32, FMB_V32T).
template <unsigned OPT, int MINB = 2>
__global__ void __launch_bounds__(V32_NT, MINB) v32t_pass_kernel(const __grid_constant__ FastArgs<float2> a, const __grid_constant__ CUtensorMap map,
                                                             const int map_col0) {
    typedef float2 C;
    extern __shared__ unsigned char v32t_smem_raw[];
    unsigned char *const base = v32t_smem_raw + ((128u - (v32p_smem_u32(v32t_smem_raw) & 127u)) & 127u);
    const unsigned base_s = v32p_smem_u32(base), bar = base_s + (unsigned)V32P_BUF;
    const int tid = threadIdx.x;
    const unsigned line0 = blockIdx.x << V32_LOGT, col = line0 >> 10, i0 = line0 & 1023u;
    if (tid == 0) {
        v32p_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        v32p_mbar_expect_tx(bar, V32P_TILE_BYTES);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            v32p_tma_box(base_s + q * 16384u, &map, bar, (int)i0, 256 * q, map_col0 + (int)col, 0x1000000000000000ull);
    }
    __syncthreads();                                   // the barrier is initialised before anybody polls it
    v32p_mbar_wait(bar, 0u);
    C *const out_col = a.out + (long long)col * a.out_cs;
    v32p_tile<OPT>(a, reinterpret_cast<C *>(base), out_col, i0, tid, 0, []() {});
}

template <unsigned OPT, int MINB> int launch_v32t_inst(const FastArgs<float2> &a, const CUtensorMap &map, int map_col0, unsigned lines, cudaStream_t st) {
    constexpr size_t smem = V32P_BUF + 128 /* alignment slack */ + 64 /* mbarrier */;
    static int attr_done = 0;
    if (!attr_done) {
        FMB_CUDA_OK(cudaFuncSetAttribute(v32t_pass_kernel<OPT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = 1;
    }
    v32t_pass_kernel<OPT, MINB><<<lines >> V32_LOGT, V32_NT, smem, st>>>(a, map, map_col0);
    FMB_LAUNCH_OK();
    return FMB_OK;
}
template <unsigned OPT> int launch_v32t_variant(const FastArgs<float2> &a, const CUtensorMap &map, int map_col0, unsigned lines, cudaStream_t st) {
#ifndef V32_LEAN
    static const int occ3 = getenv("FMB_V32T_OCC") ? atoi(getenv("FMB_V32T_OCC")) : 0;       // experiments: 3 CTAs per SM (80 registers)
    if (occ3) return launch_v32t_inst<OPT, 3>(a, map, map_col0, lines, st);
#endif
    return launch_v32t_inst<OPT, 2>(a, map, map_col0, lines, st);
}

// ---- variants (fft_engine.cu: run_v32p)
enum V32PVariant { VP_F = 0, VP_FC = 1, VP_CV_N = 2, VP_CV_M = 3, VP_CVC_N = 4, VP_CVC_M = 5, VP_K = 6, VP_KC = 7, VP_FI = 8, VP_FIC = 9 };
// in-place four-step (the intermediate lives in y itself, no ring): pass A writes its lines contiguously, y[i 1024 + k1];
// pass B fetches strided boxes of y and overwrites exactly the addresses it read, y[k2 1024 + k1]
constexpr unsigned V32P_FI_A = FO_LOAD_T | FO_TWIDDLE, V32P_FI_AC = V32P_FI_A | FO_IN_CONJ;
constexpr unsigned V32P_FI_B = FO_LOAD_T | FO_STORE_T, V32P_FI_BC = V32P_FI_B | FO_OUT_CONJ;
constexpr unsigned V32P_K_A = FO_LOAD_T | FO_STORE_T, V32P_K_AC = V32P_K_A | FO_IN_CONJ;     // Kron(Fourier, Fourier): no twiddle,
constexpr unsigned V32P_K_B = 0u, V32P_K_BC = FO_OUT_CONJ;                                    // natural order, all rows kept

template <unsigned OA, unsigned OB, unsigned OC>
int launch_v32p_variant(const V32PArgs &g, const CUtensorMap &mx, const CUtensorMap &mr, cudaStream_t st) {
    static int grid_cap = 0;
    if (!grid_cap) {
        FMB_CUDA_OK(cudaFuncSetAttribute(v32p_kernel<OA, OB, OC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V32P_SMEM));
        int per_sm = 0;
        FMB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v32p_kernel<OA, OB, OC>, V32P_NT, V32P_SMEM));
        if (per_sm < 1) { set_error("V32P kernel does not fit on an SM"); return FMB_ERR_CUDA; }
        grid_cap = device_props().sm_count;        // one CTA per SM; every CTA of the grid must be resident (polls)
    }
    unsigned grid = (unsigned)grid_cap;
    if (grid > g.total_items) grid = g.total_items;
    if (grid == 0) return FMB_OK;
    // Cooperative launch: the runtime either makes ALL CTAs of the grid resident at once or refuses the launch.  The work
    // list is only deadlock-free if they are (producers and consumers poll each other), so a context that cannot host one
    // CTA per SM right now (MPS share, green context, SMs held by another stream) must get an error here, not a hang;
    // the engine then falls back to the per-pass kernels (FMB_ERR_FALLBACK).
    void *args[3] = {const_cast<V32PArgs *>(&g), const_cast<CUtensorMap *>(&mx), const_cast<CUtensorMap *>(&mr)};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)v32p_kernel<OA, OB, OC>, dim3(grid), dim3(V32P_NT), args, V32P_SMEM, st);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources) {
        (void)cudaGetLastError();
        return FMB_ERR_FALLBACK;
    }
    if (e != cudaSuccess) { set_error("persistent kernel launch failed: %s", cudaGetErrorString(e)); return FMB_ERR_CUDA; }
    g_launches.fetch_add(1);
    return FMB_OK;
}

}  // namespace fmb
