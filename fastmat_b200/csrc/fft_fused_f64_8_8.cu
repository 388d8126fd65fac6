#include "fft_fused_inst.cuh"
namespace fmb {
int launch_fused_f64_8_8(int variant, const FusedArgs<double2> &g, cudaStream_t st) { return launch_fused_pair<double2, 8, 8>(variant, g, st); }
}  // namespace fmb
