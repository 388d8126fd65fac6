#include "common.h"
#include "fft_v32.cuh"
namespace fmb {
// returns FMB_ERR_NOTIMPL when `opt` is not one of this translation unit's variants
int launch_v32_c(unsigned opt, const FastArgs<float2> &a, unsigned tiles, cudaStream_t st) {
    switch (opt) {
        case V32_C_M: return launch_v32_variant<V32_C_M>(a, tiles, st);
        case V32_C_MP: return launch_v32_variant<V32_C_MP>(a, tiles, st);
        case V32_C_MPC: return launch_v32_variant<V32_C_MPC>(a, tiles, st);
        case V32_C_N: return launch_v32_variant<V32_C_N>(a, tiles, st);
        default: return FMB_ERR_NOTIMPL;
    }
}
}  // namespace fmb
