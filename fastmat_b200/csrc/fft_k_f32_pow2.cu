#include "fft_kernel_inst.cuh"
namespace fmb {
int launch_fft_f32_pow2(const PassParams<float2> &p, unsigned tiles, int nt, size_t smem, cudaStream_t st) {
    return launch_fft_kernel_impl<float2, true>(p, tiles, nt, smem, st);
}
}  // namespace fmb
