#include "common.h"
#include "fft_v32.cuh"
namespace fmb {
// returns FMB_ERR_NOTIMPL when `opt` is not one of this translation unit's variants
int launch_v32_b(unsigned opt, const FastArgs<float2> &a, unsigned lines, int shape, cudaStream_t st) {
    switch (opt) {
        case V32_B_F: return launch_v32_variant<V32_B_F>(a, lines, shape, st);
        case V32_B_FC: return launch_v32_variant<V32_B_FC>(a, lines, shape, st);
        case V32_K_A: return launch_v32_variant<V32_K_A>(a, lines, shape, st);
        case V32_K_AC: return launch_v32_variant<V32_K_AC>(a, lines, shape, st);
        case V32_K_B: return launch_v32_variant<V32_K_B>(a, lines, shape, st);
        case V32_K_BC: return launch_v32_variant<V32_K_BC>(a, lines, shape, st);
        case V32_B_N: return launch_v32_variant<V32_B_N>(a, lines, shape, st);
        case V32_B_NC: return launch_v32_variant<V32_B_NC>(a, lines, shape, st);
        default: return FMB_ERR_NOTIMPL;
    }
}
}  // namespace fmb
