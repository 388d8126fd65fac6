// One translation unit per (precision, log2 R1, log2 R2) instantiates the six fused pipelines.
#pragma once
#include "common.h"
#include "fft_fast_inst.cuh"
#include "fft_fused.cuh"

namespace fmb {

enum FusedVariant { FU_F = 0, FU_FC = 1, FU_CV = 2, FU_CVC = 3, FU_BL = 4, FU_BLC = 5 };

template <typename C, int L1, int L2, unsigned OA, unsigned OB, unsigned OC>
int launch_fused_variant(const FusedArgs<C> &g, cudaStream_t st, int *max_grid_out) {
    constexpr int LT1 = FusedTile<C, L1>::LOGT, LT2 = FusedTile<C, L2>::LOGT, NT = FusedTile<C, L1>::NT;
    const size_t s1 = (size_t)FastGeom<L1, LT1>::SMEM_ELEMS * sizeof(C), s2 = (size_t)FastGeom<L2, LT2>::SMEM_ELEMS * sizeof(C);
    const size_t smem = s1 > s2 ? s1 : s2;
    static int grid_cap = 0;
    if (!grid_cap) {
        FMB_CUDA_OK(cudaFuncSetAttribute(fused_kernel<C, L1, L2, OA, OB, OC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        FMB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_kernel<C, L1, L2, OA, OB, OC>, NT, smem));
        if (per_sm < 1) { set_error("fused kernel does not fit on an SM"); return FMB_ERR_CUDA; }
        grid_cap = per_sm * device_props().sm_count;          // every CTA of the grid must be resident (spin-waits)
    }
    if (max_grid_out) *max_grid_out = grid_cap;
    unsigned grid = (unsigned)grid_cap;
    if (grid > g.total_items) grid = g.total_items;
    if (grid == 0) return FMB_OK;
    fused_kernel<C, L1, L2, OA, OB, OC><<<grid, NT, smem, st>>>(g);
    FMB_LAUNCH_OK();
    return FMB_OK;
}

template <typename C, int L1, int L2> int launch_fused_pair(int variant, const FusedArgs<C> &g, cudaStream_t st) {
    constexpr unsigned CG = FO_IN_CG;
    switch (variant) {
        case FU_F: return launch_fused_variant<C, L1, L2, FV_A_F, FV_B_F | CG, 0u>(g, st, nullptr);
        case FU_FC: return launch_fused_variant<C, L1, L2, FV_A_FC, FV_B_FC | CG, 0u>(g, st, nullptr);
        case FU_CV: return launch_fused_variant<C, L1, L2, FV_A_M, FV_BM | CG, FV_C_M | CG>(g, st, nullptr);
        case FU_CVC: return launch_fused_variant<C, L1, L2, FV_A_M, FV_BMC | CG, FV_C_M | CG>(g, st, nullptr);
        case FU_BL: return launch_fused_variant<C, L1, L2, FV_A_MP, FV_BM | CG, FV_C_MP | CG>(g, st, nullptr);
        case FU_BLC: return launch_fused_variant<C, L1, L2, FV_A_MPC, FV_BMC | CG, FV_C_MPC | CG>(g, st, nullptr);
        default: set_error("fused path: unknown variant %d", variant); return FMB_ERR_NOTIMPL;
    }
}

}  // namespace fmb
