// Host-side FFT size planner, bit-identical in its decisions to the reference
// (fastmat/core/cmath.pyx:35-85 _findFFTFactors, :88-153 _findOptimalFFTSize, :156-214 _getFFTComplexity).
// The reference keeps `remaining` and `complexity` in C floats; so does this file, because those roundings decide
// padded sizes (e.g. the answer for 2^24+1 is 2^24) and hence the internal layout of Circulant / Toeplitz / Fourier.
#include <cmath>
#include <cstdint>

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

// ---- LFSRCirculant host logic (fastmat/LFSRCirculant.pyx:28-48 step functions, :196-222 order / period checks,
//      :277-314 states / vecC, :316-395 the two address sequences of _core).  Integer work, reproduced exactly.

// Fibonacci step of the generator register: feedback = parity of (state & polynomial), shifted in at bit `order`.
static inline uint32_t lfsr_gen_step(uint32_t state, uint32_t polynomial, uint32_t mask) {
    if (__builtin_parity(state & polynomial)) state |= mask;
    return state >> 1;
}

// Galois step of the tap (address) register: multiply by x modulo the polynomial.
static inline uint32_t lfsr_tap_step(uint32_t state, uint32_t polynomial, uint32_t mask) {
    state <<= 1;
    if (state & mask) state ^= (polynomial | mask);
    return state;
}

int lfsr_order(uint32_t polynomial) {
    int order = 0;
    uint32_t mask = 1;
    while ((~mask & polynomial) > mask) { mask <<= 1; ++order; }
    return order;
}

// period of the register, or -1 (order outside 1..31), -2 (zero start), -3 (sequence never returns to start)
int64_t lfsr_period(uint32_t polynomial, uint32_t start) {
    const int order = lfsr_order(polynomial);
    if (order > 31 || order < 1) return -1;
    const uint32_t mask = 1u << order;
    start &= (mask - 1);
    if (start == 0) return -2;
    uint32_t state = lfsr_gen_step(start, polynomial, mask);
    int64_t period = 1;
    while (state != start) {
        state = lfsr_gen_step(state, polynomial, mask);
        ++period;
        if (period >= (int64_t)mask || state == 0) return -3;
    }
    return period;
}

// n successive generator states from `start`, tap states from 1, and the +1/-1 output sequence (any may be null)
void lfsr_sequences(uint32_t polynomial, uint32_t start, int64_t n, uint32_t *gen_states, uint32_t *tap_states,
                    int8_t *vec_c) {
    const uint32_t mask = 1u << lfsr_order(polynomial);
    uint32_t g = start & (mask - 1), t = 1;
    for (int64_t i = 0; i < n; ++i) {
        if (gen_states) gen_states[i] = g & (mask - 1);
        if (tap_states) tap_states[i] = t;
        if (vec_c) vec_c[i] = (g & 1) ? -1 : 1;
        g = lfsr_gen_step(g, polynomial, mask);
        t = lfsr_tap_step(t, polynomial, mask);
    }
}

}  // namespace fmb
