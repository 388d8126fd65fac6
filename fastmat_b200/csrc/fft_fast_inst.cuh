// One translation unit per (precision, log2 R) instantiates the pass variants of the fast path.
#pragma once
#include "common.h"
#include "fft_fast.cuh"

namespace fmb {

// the pass variants the engine uses (see fft_engine.cu: run_fast)
constexpr unsigned FV_A_F = FO_LOAD_T | FO_TWIDDLE;                                   // first pass, plain
constexpr unsigned FV_A_FC = FV_A_F | FO_IN_CONJ;                                     // first pass of conj(F conj x)
constexpr unsigned FV_A_M = FV_A_F | FO_IN_MASK;                                      // first pass, zero-padded input
constexpr unsigned FV_A_MP = FV_A_M | FO_PRE;                                         // ... with pre-multiply (chirp)
constexpr unsigned FV_A_MPC = FV_A_MP | FO_PRE_CONJ;
constexpr unsigned FV_B_F = FO_LOAD_T | FO_STORE_T | FO_OUT_MASK;                     // second pass of a plain transform
constexpr unsigned FV_B_FC = FV_B_F | FO_OUT_CONJ;
constexpr unsigned FV_BM = FO_LOAD_T | FO_STORE_T | FO_TWO_FFTS | FO_TWIDDLE;         // middle pass of a convolution
constexpr unsigned FV_BMC = FV_BM | FO_MID_CONJ;
constexpr unsigned FV_C_M = FO_STORE_T | FO_OUT_CONJ | FO_OUT_MASK;                   // last pass of a convolution
constexpr unsigned FV_C_MP = FV_C_M | FO_POST;
constexpr unsigned FV_C_MPC = FV_C_MP | FO_POST_CONJ;
constexpr unsigned FV_K_AC = FV_B_F | FO_IN_CONJ;                                    // Kron: first pass of the backward
constexpr unsigned FV_K_B = FO_OUT_MASK;                                             // Kron: second pass (rows contiguous)
constexpr unsigned FV_K_BC = FO_OUT_MASK | FO_OUT_CONJ;
// whole transforms in one kernel (L <= 4096: a line is a column of the operand, see fft_engine.cu: run_single_fast)
constexpr unsigned FV_1_FC = FO_IN_CONJ | FO_OUT_CONJ | FO_OUT_MASK;                  // backward Fourier (forward: FV_K_B)
constexpr unsigned FV_1_M = FO_TWO_FFTS | FO_IN_MASK | FO_OUT_CONJ | FO_OUT_MASK;     // Circulant / Toeplitz on chip
constexpr unsigned FV_1_MC = FV_1_M | FO_MID_CONJ;
constexpr unsigned FV_1_MP = FV_1_M | FO_PRE | FO_POST;                                // ... chirp-z on chip: pre- and post-multiply
constexpr unsigned FV_1_MPC = FV_1_MP | FO_MID_CONJ | FO_PRE_CONJ | FO_POST_CONJ;
constexpr unsigned FV_1T = FO_LOAD_T | FO_STORE_T;                                    // the same for row-major operands: the
constexpr unsigned FV_1T_FC = FV_1_FC | FV_1T;                                        // T columns of a tile are contiguous
constexpr unsigned FV_1T_M = FV_1_M | FV_1T;                                          // (forward Fourier: FV_B_F)
constexpr unsigned FV_1T_MC = FV_1_MC | FV_1T;
constexpr unsigned FV_1T_MP = FV_1_MP | FV_1T;
constexpr unsigned FV_1T_MPC = FV_1_MPC | FV_1T;

template <typename C, int LOGR, unsigned OPT>
int launch_fast_variant(const FastArgs<C> &a, unsigned tiles, cudaStream_t st) {
    constexpr int LOGT = FastTile<LOGR>::LOGT - (sizeof(C) == 16 ? 1 : 0);              // complex128: half the lines per tile
    typedef FastGeom<LOGR, LOGT> G;
    const size_t smem = (size_t)G::SMEM_ELEMS * sizeof(C);
    static int attr_done = 0;
    if (!attr_done) {
        FMB_CUDA_OK(cudaFuncSetAttribute(fast_pass_kernel<C, LOGR, LOGT, OPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = 1;
    }
    fast_pass_kernel<C, LOGR, LOGT, OPT><<<tiles, G::NT, smem, st>>>(a);
    FMB_LAUNCH_OK();
    return FMB_OK;
}

template <typename C, int LOGR> int launch_fast_logr(unsigned opt, const FastArgs<C> &a, unsigned tiles, cudaStream_t st) {
    switch (opt) {
        case FV_A_F: return launch_fast_variant<C, LOGR, FV_A_F>(a, tiles, st);
        case FV_A_FC: return launch_fast_variant<C, LOGR, FV_A_FC>(a, tiles, st);
        case FV_A_M: return launch_fast_variant<C, LOGR, FV_A_M>(a, tiles, st);
        case FV_A_MP: return launch_fast_variant<C, LOGR, FV_A_MP>(a, tiles, st);
        case FV_A_MPC: return launch_fast_variant<C, LOGR, FV_A_MPC>(a, tiles, st);
        case FV_B_F: return launch_fast_variant<C, LOGR, FV_B_F>(a, tiles, st);
        case FV_B_FC: return launch_fast_variant<C, LOGR, FV_B_FC>(a, tiles, st);
        case FV_BM: return launch_fast_variant<C, LOGR, FV_BM>(a, tiles, st);
        case FV_BMC: return launch_fast_variant<C, LOGR, FV_BMC>(a, tiles, st);
        case FV_C_M: return launch_fast_variant<C, LOGR, FV_C_M>(a, tiles, st);
        case FV_C_MP: return launch_fast_variant<C, LOGR, FV_C_MP>(a, tiles, st);
        case FV_C_MPC: return launch_fast_variant<C, LOGR, FV_C_MPC>(a, tiles, st);
        case FV_K_AC: return launch_fast_variant<C, LOGR, FV_K_AC>(a, tiles, st);
        case FV_K_B: return launch_fast_variant<C, LOGR, FV_K_B>(a, tiles, st);
        case FV_K_BC: return launch_fast_variant<C, LOGR, FV_K_BC>(a, tiles, st);
        case FV_1_FC: return launch_fast_variant<C, LOGR, FV_1_FC>(a, tiles, st);
        case FV_1_M: return launch_fast_variant<C, LOGR, FV_1_M>(a, tiles, st);
        case FV_1_MC: return launch_fast_variant<C, LOGR, FV_1_MC>(a, tiles, st);
        case FV_1T_FC: return launch_fast_variant<C, LOGR, FV_1T_FC>(a, tiles, st);
        case FV_1T_M: return launch_fast_variant<C, LOGR, FV_1T_M>(a, tiles, st);
        case FV_1T_MC: return launch_fast_variant<C, LOGR, FV_1T_MC>(a, tiles, st);
        case FV_1_MP: return launch_fast_variant<C, LOGR, FV_1_MP>(a, tiles, st);
        case FV_1_MPC: return launch_fast_variant<C, LOGR, FV_1_MPC>(a, tiles, st);
        case FV_1T_MP: return launch_fast_variant<C, LOGR, FV_1T_MP>(a, tiles, st);
        case FV_1T_MPC: return launch_fast_variant<C, LOGR, FV_1T_MPC>(a, tiles, st);
        default: set_error("fast path: unknown pass variant %u", opt); return FMB_ERR_NOTIMPL;
    }
}

}  // namespace fmb
