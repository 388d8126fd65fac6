#include "common.h"
#include "fft_v32.cuh"
namespace fmb {
// returns FMB_ERR_NOTIMPL when `opt` is not one of this translation unit's variants
int launch_v32_a(unsigned opt, const FastArgs<float2> &a, unsigned lines, int shape, cudaStream_t st) {
    switch (opt) {
        case V32_A_F: return launch_v32_variant<V32_A_F>(a, lines, shape, st);
        case V32_A_FC: return launch_v32_variant<V32_A_FC>(a, lines, shape, st);
        case V32_A_M: return launch_v32_variant<V32_A_M>(a, lines, shape, st);
        case V32_A_MP: return launch_v32_variant<V32_A_MP>(a, lines, shape, st);
        case V32_A_MPC: return launch_v32_variant<V32_A_MPC>(a, lines, shape, st);
        case V32_A_H: return launch_v32_variant<V32_A_H>(a, lines, shape, st);
        default: return FMB_ERR_NOTIMPL;
    }
}
}  // namespace fmb
