#include "fft_fused_inst.cuh"
namespace fmb {
int launch_fused_f32_10_11(int variant, const FusedArgs<float2> &g, cudaStream_t st) { return launch_fused_pair<float2, 10, 11>(variant, g, st); }
}  // namespace fmb
