#include "fft_fast_inst.cuh"
namespace fmb {
int launch_fast_f64_L6(unsigned opt, const FastArgs<double2> &a, unsigned tiles, cudaStream_t st) {
    return launch_fast_logr<double2, 6>(opt, a, tiles, st);
}
}  // namespace fmb
