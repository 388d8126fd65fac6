#include "common.h"
#include "fft_v32p.cuh"
namespace fmb {
// TMA-fed strided passes of the pipelined-slab schedule; returns FMB_ERR_NOTIMPL when `opt` is not one of the variants
int launch_v32t(unsigned opt, const FastArgs<float2> &a, const CUtensorMap &map, int map_col0, unsigned lines, cudaStream_t st) {
    switch (opt) {
        case V32_A_F: return launch_v32t_variant<V32_A_F>(a, map, map_col0, lines, st);
        case V32_A_FC: return launch_v32t_variant<V32_A_FC>(a, map, map_col0, lines, st);
        case V32_C_N: return launch_v32t_variant<V32_C_N>(a, map, map_col0, lines, st);
        case V32_C_M: return launch_v32t_variant<V32_C_M>(a, map, map_col0, lines, st);
        case (V32_C_N | V32_C_TW): return launch_v32t_variant<(V32_C_N | V32_C_TW)>(a, map, map_col0, lines, st);
        case (V32_C_M | V32_C_TW): return launch_v32t_variant<(V32_C_M | V32_C_TW)>(a, map, map_col0, lines, st);
        default: return FMB_ERR_NOTIMPL;
    }
}
}  // namespace fmb
