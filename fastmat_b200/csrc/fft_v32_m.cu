#include "common.h"
#include "fft_v32.cuh"
namespace fmb {
// returns FMB_ERR_NOTIMPL when `opt` is not one of this translation unit's variants
int launch_v32_m(unsigned opt, const FastArgs<float2> &a, unsigned lines, int shape, cudaStream_t st) {
    switch (opt) {
        case V32_BM: return launch_v32_variant<V32_BM>(a, lines, shape, st);
        case V32_BMC: return launch_v32_variant<V32_BMC>(a, lines, shape, st);
        case V32_BM_N: return launch_v32_variant<V32_BM_N>(a, lines, shape, st);
        case V32_BMC_N: return launch_v32_variant<V32_BMC_N>(a, lines, shape, st);
        default: return FMB_ERR_NOTIMPL;
    }
}
}  // namespace fmb
