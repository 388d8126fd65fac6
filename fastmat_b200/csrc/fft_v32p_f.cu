#include "common.h"
#include "fft_v32p.cuh"
namespace fmb {
// plain transforms (1-D four-step and the 2-D transform of Kron(Fourier(1024), Fourier(1024))); returns FMB_ERR_NOTIMPL when `variant` is not one of this translation unit's
int launch_v32p_f(int variant, const V32PArgs &g, const CUtensorMap &mx, const CUtensorMap &mr, cudaStream_t st) {
    switch (variant) {
        case VP_F: return launch_v32p_variant<V32_A_F, V32_B_N, 0u>(g, mx, mr, st);
        case VP_FC: return launch_v32p_variant<V32_A_FC, V32_B_NC, 0u>(g, mx, mr, st);
        case VP_K: return launch_v32p_variant<V32P_K_A, V32P_K_B, 0u>(g, mx, mr, st);
        case VP_KC: return launch_v32p_variant<V32P_K_AC, V32P_K_BC, 0u>(g, mx, mr, st);
        case VP_FI: return launch_v32p_variant<V32P_FI_A, V32P_FI_B, 0u>(g, mx, mr, st);
        case VP_FIC: return launch_v32p_variant<V32P_FI_AC, V32P_FI_BC, 0u>(g, mx, mr, st);
        default: return FMB_ERR_NOTIMPL;
    }
}
}  // namespace fmb
