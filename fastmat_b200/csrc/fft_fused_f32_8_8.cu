#include "fft_fused_inst.cuh"
namespace fmb {
int launch_fused_f32_8_8(int variant, const FusedArgs<float2> &g, cudaStream_t st) { return launch_fused_pair<float2, 8, 8>(variant, g, st); }
}  // namespace fmb
