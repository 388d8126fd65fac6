#include "common.h"
#include "fft_v32p.cuh"
namespace fmb {
// convolutions, forward; returns FMB_ERR_NOTIMPL when `variant` is not one of this translation unit's
int launch_v32p_c0(int variant, const V32PArgs &g, const CUtensorMap &mx, const CUtensorMap &mr, cudaStream_t st) {
    switch (variant) {
        case VP_CV_N: return launch_v32p_variant<V32_A_F, V32_BM, V32_C_N>(g, mx, mr, st);
        case VP_CV_M: return launch_v32p_variant<V32_A_F, V32_BM, V32_C_M>(g, mx, mr, st);
        default: return FMB_ERR_NOTIMPL;
    }
}
}  // namespace fmb
