// Kernel wrapper + launcher for one (complex type, POW2) instantiation of the FFT pass; each instantiation lives in
// its own translation unit (fft_k_*.cu) so that the four heavy device compiles run in parallel.
#pragma once
#include "common.h"
#include "fft_pass.cuh"

namespace fmb {

struct DevSync {
    FMB_HD void operator()() const {
#ifdef __CUDA_ARCH__
        __syncthreads();
#endif
    }
};

template <typename C, bool POW2>
__global__ void __launch_bounds__(FMB_MAX_NT, 1) fft_pass_kernel(const __grid_constant__ PassParams<C> p) {
    extern __shared__ __align__(16) unsigned char fmb_smem_raw[];
    DevSync sync;
    pass_body<C, POW2, DevSync>(p, (long long)blockIdx.x, (int)threadIdx.x, (int)blockDim.x, reinterpret_cast<C *>(fmb_smem_raw), sync);
}

template <typename C, bool POW2>
int launch_fft_kernel_impl(const PassParams<C> &p, unsigned tiles, int nt, size_t smem, cudaStream_t st) {
    static int attr_done = 0;                       // benign race: setting the attribute is idempotent
    if (!attr_done) {
        FMB_CUDA_OK(cudaFuncSetAttribute(fft_pass_kernel<C, POW2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)device_props().smem_optin));
        attr_done = 1;
    }
    fft_pass_kernel<C, POW2><<<tiles, nt, smem, st>>>(p);
    FMB_LAUNCH_OK();
    return FMB_OK;
}

}  // namespace fmb
