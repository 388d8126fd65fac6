// Host-side FFT size planner, bit-identical in its decisions to the reference
// (fastmat/core/cmath.pyx:35-85 _findFFTFactors, :88-153 _findOptimalFFTSize, :156-214 _getFFTComplexity).
// The reference keeps `remaining` and `complexity` in C floats; so does this file, because those roundings decide
// padded sizes (e.g. the answer for 2^24+1 is 2^24) and hence the internal layout of Circulant / Toeplitz / Fourier.
#include <cmath>

#include "common.h"

namespace fmb {

// state = (complexity << 16) + length, greedy search over stage sizes <= max_factor
static int search_factors(int target_length, int max_factor, int state, int best_state) {
    for (int ff = max_factor; ff > 0; --ff) {
        const int length = (state & 0xFFFF) * ff;
        const int complexity = (state >> 16) + ff + 1;
        const int next = (complexity << 16) + length;
        if (next <= best_state && length < target_length) {
            best_state = search_factors(target_length, ff, next, best_state);
        } else if (next < best_state) {
            best_state = next;
        }
    }
    return best_state;
}

int64_t find_optimal_fft_size(int64_t order, int max_stage) {
    int64_t padded = 1;
    float remaining = (float)order;
    while (remaining > 64) {          // peel radix-4 stages until at most 64 is left
        padded *= 4;
        remaining /= 4;
    }
    const int x = (int)std::ceil(remaining);
    if (x != 1) {
        const int start_best = ((3 * (4 + 1)) << 16) + 64;      // three radix-4 stages reach 64
        padded *= search_factors(x, max_stage, 1, start_best) & 0xFFFF;
    }
    return padded;
}

float fft_complexity(int64_t n) {
    float complexity = 0;
    const float float_n = (float)n;
    int64_t nn = n;
    while (nn % 4 == 0) { complexity += 4 + 1; nn /= 4; }
    if (nn > 1 && (nn & 1) == 0) { complexity += 2 + 1; nn /= 2; }
    int64_t ii = 3;
    while (nn > 1 && ii * ii < nn) {  // strict '<': a leftover square p*p is charged as a single stage
        if (nn % ii == 0) { complexity += (float)(ii + 1); nn /= ii; }
        else ii += 2;
    }
    if (nn > 1) complexity += (float)(nn + 1);
    return float_n * (complexity + 1);
}

}  // namespace fmb
