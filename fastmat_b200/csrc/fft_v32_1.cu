#include "common.h"
#include "fft_v32.cuh"
namespace fmb {
// whole convolutions of FFT length 1024 in one kernel (Circulant(1024), Toeplitz padded to 1024); returns FMB_ERR_NOTIMPL
// when `opt` is not one of this translation unit's variants
int launch_v32_1(unsigned opt, const FastArgs<float2> &a, unsigned lines, int shape, cudaStream_t st) {
    switch (opt) {
        case V32_1M: return launch_v32_variant<V32_1M>(a, lines, shape, st);
        case V32_1MC: return launch_v32_variant<V32_1MC>(a, lines, shape, st);
        case V32_1H: return launch_v32_variant<V32_1H>(a, lines, shape, st);
        case V32_1HC: return launch_v32_variant<V32_1HC>(a, lines, shape, st);
        case V32_1K: return launch_v32_variant<V32_1K>(a, lines, shape, st);
        case V32_1KC: return launch_v32_variant<V32_1KC>(a, lines, shape, st);
        default: return FMB_ERR_NOTIMPL;
    }
}
}  // namespace fmb
