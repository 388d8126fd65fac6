#include "common.h"
#include "fft_v32.cuh"
namespace fmb {
// returns FMB_ERR_NOTIMPL when `opt` is not one of this translation unit's variants
int launch_v32_c(unsigned opt, const FastArgs<float2> &a, unsigned lines, int shape, cudaStream_t st) {
    switch (opt) {
        case V32_C_M: return launch_v32_variant<V32_C_M>(a, lines, shape, st);
        case V32_C_MP: return launch_v32_variant<V32_C_MP>(a, lines, shape, st);
        case V32_C_MPC: return launch_v32_variant<V32_C_MPC>(a, lines, shape, st);
        case V32_C_N: return launch_v32_variant<V32_C_N>(a, lines, shape, st);
        case (V32_C_M | V32_C_TW): return launch_v32_variant<(V32_C_M | V32_C_TW)>(a, lines, shape, st);
        case (V32_C_MP | V32_C_TW): return launch_v32_variant<(V32_C_MP | V32_C_TW)>(a, lines, shape, st);
        case (V32_C_MPC | V32_C_TW): return launch_v32_variant<(V32_C_MPC | V32_C_TW)>(a, lines, shape, st);
        case (V32_C_N | V32_C_TW): return launch_v32_variant<(V32_C_N | V32_C_TW)>(a, lines, shape, st);
        case V32_C_H: return launch_v32_variant<V32_C_H>(a, lines, shape, st);
        case (V32_C_H | V32_C_TW): return launch_v32_variant<(V32_C_H | V32_C_TW)>(a, lines, shape, st);
        default: return FMB_ERR_NOTIMPL;
    }
}
}  // namespace fmb
