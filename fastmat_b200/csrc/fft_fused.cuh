// Fused persistent kernel of the fast path: ALL passes of a large transform / convolution in ONE launch.
//
// The column batch is cut into slabs of a few columns.  A slab's intermediate (its four-step "tmp" array) lives in a
// small ring of slots that stays resident in the 126 MB L2, so HBM sees each input element once and each output
// element once.  One CTA per SM-slot runs a loop over a global work list ordered  A(s), B(s-D), C(s-2D)  for
// s = 0, 1, ...; a tile of pass p of slab s may start when all tiles of pass p-1 of slab s have signalled completion
// (global counters, release/acquire at GPU scope), and pass A of slab s may overwrite the ring slot once the last pass of
// slab s - NSLOT has finished.  Every dependency points to a work item with a smaller index, all CTAs of the grid are
// co-resident, hence the spin-waits cannot deadlock.
#pragma once
#include "fft_fast.cuh"

namespace fmb {

constexpr int FUSED_MAX_PASS = 3;

template <typename C> struct FusedArgs {
    FastArgs<C> pass[FUSED_MAX_PASS];   // in / out of pass 0 and of the last pass are set for slab 0; tmp for slot 0
    int npass;
    int ncols, slab_cols, nslabs;
    unsigned tiles[FUSED_MAX_PASS];     // tiles per (full) slab and pass
    unsigned items_per_step;            // sum of tiles[]
    int delay, nslot;
    long long slot_stride;              // elements between ring slots
    long long x_slab_stride, y_slab_stride;
    unsigned *work_counter;
    unsigned *done;                     // [npass][nslabs]
    unsigned total_items;
};

// Flag protocol WITHOUT L1 invalidation.  A gpu-scope acquire load or __threadfence() makes ptxas emit CCTL.IVALL, which
// throws away the SM's whole L1 (twiddle tables included) once per tile.  It is not needed here: the data guarded by the
// flags (the ring slots) is only ever read with ld.global.cg, i.e. straight from L2, never from L1.  So the consumer
// polls with a relaxed load and the producer publishes with a release reduction (MEMBAR.ALL.GPU + RED, no CCTL).
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename C, int LOGR> struct FusedTile {      // 8192 elements per tile (complex128: 4096), i.e. one CTA shape for all passes
    static constexpr int LOGT = 13 - LOGR - (sizeof(C) == 16 ? 1 : 0);
    static constexpr int NT = FastGeom<LOGR, LOGT>::NT;
};

template <typename C, int LOGR1, int LOGR2, unsigned OPT_A, unsigned OPT_B, unsigned OPT_C>
__global__ void __launch_bounds__(FusedTile<C, LOGR1>::NT, (FusedTile<C, LOGR1>::NT >= 512 ? 2 : 3))
fused_kernel(const __grid_constant__ FusedArgs<C> g) {
    constexpr int LT1 = FusedTile<C, LOGR1>::LOGT, LT2 = FusedTile<C, LOGR2>::LOGT;
    static_assert(FastGeom<LOGR1, LT1>::NT == FastGeom<LOGR2, LT2>::NT, "all passes of a fused kernel share the CTA shape");
    __shared__ unsigned s_item;
    constexpr int last = (OPT_C == 0) ? 1 : 2;            // two passes (plain transform) or three (convolution)
    for (;;) {
        __syncthreads();                                  // previous tile done with shared memory and s_item
        if (threadIdx.x == 0) s_item = atomicAdd(g.work_counter, 1u);
        __syncthreads();
        const unsigned w = s_item;
        if (w >= g.total_items) break;
        const unsigned step = w / g.items_per_step, r = w - step * g.items_per_step;
        int p;
        unsigned tile;
        if (r < g.tiles[0]) { p = 0; tile = r; }
        else if (r < g.tiles[0] + g.tiles[1]) { p = 1; tile = r - g.tiles[0]; }
        else { p = 2; tile = r - g.tiles[0] - g.tiles[1]; }
        const int slab = (int)step - p * g.delay;
        if (slab < 0 || slab >= g.nslabs) continue;       // uniform over the CTA
        const int slot = slab % g.nslot;
        // ---- wait for the producer of this tile's input (and, for pass A, for the ring slot to be free)
        if (threadIdx.x == 0) {
            const unsigned *flag = nullptr;
            unsigned target = 0;
            if (p == 0) {
                const int prev = slab - g.nslot;
                if (prev >= 0) { flag = g.done + last * g.nslabs + prev; target = g.tiles[last]; }
            } else {
                flag = g.done + (p - 1) * g.nslabs + slab;
                target = g.tiles[p - 1];
            }
            if (flag) while (ld_relaxed_gpu(flag) < target) __nanosleep(64);
        }
        __syncthreads();
        // ---- run the tile (columns beyond the batch in a ragged last slab are skipped but still signalled)
        const int cols_here = min(g.slab_cols, g.ncols - slab * g.slab_cols);
        const long long tmp_off = (long long)slot * g.slot_stride;
        if (p == 0) {
            const unsigned col = (tile << LT1) >> g.pass[0].logI;
            if ((int)col < cols_here) fast_pass_call<C, LOGR1, LT1, OPT_A>(g.pass[0], tile, (long long)slab * g.x_slab_stride, tmp_off);
        } else if (p == last) {
            if constexpr (last == 1) {
                const unsigned col = (tile << LT2) >> g.pass[1].logI;
                if ((int)col < cols_here) fast_pass_call<C, LOGR2, LT2, OPT_B>(g.pass[1], tile, tmp_off, (long long)slab * g.y_slab_stride);
            } else {
                const unsigned col = (tile << LT1) >> g.pass[2].logI;
                if ((int)col < cols_here) fast_pass_call<C, LOGR1, LT1, OPT_C>(g.pass[2], tile, tmp_off, (long long)slab * g.y_slab_stride);
            }
        } else {
            const unsigned col = (tile << LT2) >> g.pass[1].logI;
            if ((int)col < cols_here) fast_pass_call<C, LOGR2, LT2, OPT_B>(g.pass[1], tile, tmp_off, tmp_off);
        }
        // ---- signal completion: every thread's global stores precede the barrier, thread 0 publishes at GPU scope
        __syncthreads();
        if (threadIdx.x == 0) red_release_gpu(g.done + p * g.nslabs + slab, 1u);
    }
}

}  // namespace fmb
