#include "fft_fast_inst.cuh"
namespace fmb {
int launch_fast_f32_L6(unsigned opt, const FastArgs<float2> &a, unsigned tiles, cudaStream_t st) {
    return launch_fast_logr<float2, 6>(opt, a, tiles, st);
}
}  // namespace fmb
