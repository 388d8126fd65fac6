// Complex helpers and small-radix DFT butterflies shared by the FFT / FWHT kernels.
// Everything is __host__ __device__ so that tests/emul can run the exact kernel index logic on the CPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FMB_HD __host__ __device__ __forceinline__

namespace fmb {

template <typename S> struct cplx_of;
template <> struct cplx_of<float> { typedef float2 type; };
template <> struct cplx_of<double> { typedef double2 type; };
template <typename C> struct real_of;
template <> struct real_of<float2> { typedef float type; };
template <> struct real_of<double2> { typedef double type; };

template <typename C> FMB_HD C mk(typename real_of<C>::type x, typename real_of<C>::type y) { C r; r.x = x; r.y = y; return r; }
// complex add / sub.  On sm_100 a float2 is added with ONE packed instruction (FADD2, PTX add.rn.f32x2): two independent
// IEEE additions, bit-identical to the scalar pair, at half the issue slots (measured: tools/ubench_fp32.cu).
template <typename C> FMB_HD C cadd(C a, C b) { return mk<C>(a.x + b.x, a.y + b.y); }
template <typename C> FMB_HD C csub(C a, C b) { return mk<C>(a.x - b.x, a.y - b.y); }
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(FMB_NO_PACKED_F32)
__device__ __forceinline__ unsigned long long fmb_pk(float2 v) { return *reinterpret_cast<unsigned long long *>(&v); }
__device__ __forceinline__ float2 fmb_upk(unsigned long long v) { return *reinterpret_cast<float2 *>(&v); }
template <> __device__ __forceinline__ float2 cadd<float2>(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(fmb_pk(a)), "l"(fmb_pk(b)));
    return fmb_upk(r);
}
template <> __device__ __forceinline__ float2 csub<float2>(float2 a, float2 b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(fmb_pk(a)), "l"(fmb_pk(b)));
    return fmb_upk(r);
}
#endif
template <typename C> FMB_HD C cmul(C a, C b) { return mk<C>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
template <typename C> FMB_HD C cmulc(C a, C b) { return mk<C>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
template <typename C> FMB_HD C cconj(C a) { return mk<C>(a.x, -a.y); }
template <typename C> FMB_HD C cmul_mi(C a) { return mk<C>(a.y, -a.x); }   // a * (-i)
template <typename C> FMB_HD C cmul_pi(C a) { return mk<C>(-a.y, a.x); }   // a * (+i)
template <typename C> FMB_HD C cscale(C a, typename real_of<C>::type s) { return mk<C>(a.x * s, a.y * s); }

// ---- fused butterflies.  The FFT kernels are bound by FP32 issue, so a twiddle multiplication that
// feeds a radix-2 butterfly is not done as "multiply, then add and subtract" (4 + 4 operations per complex pair) but
// folded into fused multiply-adds:  t0 = a + w z  is two FMAs per component, and  t1 = a - w z = 2 a - t0  one more:
// 6 operations.  Two twiddled inputs  t2 = w1 b + w3 d,  dd = w1 b - w3 d = 2 (w1 b) - t2  cost 10 instead of 12.
FMB_HD float fmb_fma(float a, float b, float c) { return fmaf(a, b, c); }
FMB_HD double fmb_fma(double a, double b, double c) { return fma(a, b, c); }
template <typename C> FMB_HD void bfly_w(C a, C z, typename real_of<C>::type wr, typename real_of<C>::type wi, C &t0, C &t1) {
    typedef typename real_of<C>::type S;
    t0.x = fmb_fma(wr, z.x, fmb_fma(-wi, z.y, a.x));
    t0.y = fmb_fma(wr, z.y, fmb_fma(wi, z.x, a.y));
    t1.x = fmb_fma((S)2, a.x, -t0.x);
    t1.y = fmb_fma((S)2, a.y, -t0.y);
}
// t0 = a + w z only (the other half of the butterfly is not needed)
template <typename C> FMB_HD C add_w(C a, C z, typename real_of<C>::type wr, typename real_of<C>::type wi) {
    return mk<C>(fmb_fma(wr, z.x, fmb_fma(-wi, z.y, a.x)), fmb_fma(wr, z.y, fmb_fma(wi, z.x, a.y)));
}
template <typename C>
FMB_HD void bfly_ww(C b, C d, typename real_of<C>::type w1r, typename real_of<C>::type w1i, typename real_of<C>::type w3r,
                    typename real_of<C>::type w3i, C &t2, C &dd) {
    typedef typename real_of<C>::type S;
    const C p = mk<C>(fmb_fma(w1r, b.x, -w1i * b.y), fmb_fma(w1r, b.y, w1i * b.x));
    t2.x = fmb_fma(w3r, d.x, fmb_fma(-w3i, d.y, p.x));
    t2.y = fmb_fma(w3r, d.y, fmb_fma(w3i, d.x, p.y));
    dd.x = fmb_fma((S)2, p.x, -t2.x);
    dd.y = fmb_fma((S)2, p.y, -t2.y);
}
// radix-4 butterfly whose inputs 1, 2, 3 still carry the twiddles w1, w2, w3 (input 0 none):
//   X[k] = v0 + (-i)^k w1 v1 + (-1)^k w2 v2 + (i)^k w3 v3      -> v0, v1, v2, v3 = X[0], X[1], X[2], X[3]
template <typename C>
FMB_HD void dft4_tw(C &v0, C &v1, C &v2, C &v3, typename real_of<C>::type w1r, typename real_of<C>::type w1i,
                    typename real_of<C>::type w2r, typename real_of<C>::type w2i, typename real_of<C>::type w3r,
                    typename real_of<C>::type w3i) {
    C t0, t1, t2, d;
    bfly_w(v0, v2, w2r, w2i, t0, t1);
    bfly_ww(v1, v3, w1r, w1i, w3r, w3i, t2, d);
    v0 = cadd(t0, t2); v2 = csub(t0, t2);
    v1 = mk<C>(t1.x + d.y, t1.y - d.x);
    v3 = mk<C>(t1.x - d.y, t1.y + d.x);
}
// ... the same with w2 = -i (no multiplication for input 2)
template <typename C>
FMB_HD void dft4_tw_mi(C &v0, C &v1, C &v2, C &v3, typename real_of<C>::type w1r, typename real_of<C>::type w1i,
                       typename real_of<C>::type w3r, typename real_of<C>::type w3i) {
    const C t0 = mk<C>(v0.x + v2.y, v0.y - v2.x), t1 = mk<C>(v0.x - v2.y, v0.y + v2.x);
    C t2, d;
    bfly_ww(v1, v3, w1r, w1i, w3r, w3i, t2, d);
    v0 = cadd(t0, t2); v2 = csub(t0, t2);
    v1 = mk<C>(t1.x + d.y, t1.y - d.x);
    v3 = mk<C>(t1.x - d.y, t1.y + d.x);
}

// ---- forward DFTs (kernel e^{-2 pi i rk/P}), in place.  The result X[q] ends up at array position
// ---- outpos<P>(q) (digit-reversed for the composite radices) so that no register shuffling is needed.
template <int P> FMB_HD constexpr int outpos(int q) { return q; }
template <> FMB_HD constexpr int outpos<8>(int q) { return ((q & 3) << 1) | (q >> 2); }     // q = k0 + 4*k1 at 2*k0 + k1
template <> FMB_HD constexpr int outpos<16>(int q) { return ((q & 3) << 2) | (q >> 2); }    // q = k0 + 4*k1 at 4*k0 + k1

template <typename C> FMB_HD void dft2(C &a, C &b) { C t = a; a = cadd(t, b); b = csub(t, b); }

template <typename C> FMB_HD void dft4(C &v0, C &v1, C &v2, C &v3) {
    C t0 = cadd(v0, v2), t1 = csub(v0, v2), t2 = cadd(v1, v3), d = csub(v1, v3);
    v0 = cadd(t0, t2); v2 = csub(t0, t2);
    // t1 -/+ i d with scalar operations: the (y, -x) rotation costs no instruction this way
    v1 = mk<C>(t1.x + d.y, t1.y - d.x);
    v3 = mk<C>(t1.x - d.y, t1.y + d.x);
}

template <typename C> FMB_HD void dft3(C &v0, C &v1, C &v2) {
    typedef typename real_of<C>::type S;
    const S s = (S)0.86602540378443864676372317075294;
    C t1 = cadd(v1, v2);
    C t2 = mk<C>(v0.x - (S)0.5 * t1.x, v0.y - (S)0.5 * t1.y);
    C t3 = cmul_mi(cscale(csub(v1, v2), s));
    v0 = cadd(v0, t1); v1 = cadd(t2, t3); v2 = csub(t2, t3);
}

// radix 8: r = 2*r1 + r0 (r0 in 0..1), q = k0 + 4*k1 (k0 in 0..3, k1 in 0..1):
//   X[k0 + 4 k1] = sum_{r0} W8^{r0 k0} (-1)^{r0 k1} [ sum_{r1} v[2 r1 + r0] W4^{r1 k0} ]
template <typename C> FMB_HD void dft8(C *v) {
    typedef typename real_of<C>::type S;
    const S h = (S)0.70710678118654752440084436210485;
    dft4(v[0], v[2], v[4], v[6]);       // y[r0=0][k0] at v[2*k0]
    dft4(v[1], v[3], v[5], v[7]);       // y[r0=1][k0] at v[2*k0+1]
    // twiddle y[1][k0] by W8^{k0}
    C a = v[3]; v[3] = mk<C>((a.x + a.y) * h, (a.y - a.x) * h);            // * (1 - i)/sqrt2
    v[5] = cmul_mi(v[5]);                                                   // * -i
    a = v[7]; v[7] = mk<C>((a.y - a.x) * h, -(a.x + a.y) * h);             // * (-1 - i)/sqrt2
    dft2(v[0], v[1]); dft2(v[2], v[3]); dft2(v[4], v[5]); dft2(v[6], v[7]); // X[k0 + 4 k1] at v[2*k0 + k1]
}

// radix 16: r = 4*r1 + r0, q = k0 + 4*k1; X[k0 + 4 k1] = sum_{r0} W16^{r0 k0} W4^{r0 k1} [sum_{r1} v[4 r1 + r0] W4^{r1 k0}]
// The twiddles W16^{r0 k0} between the two levels are folded into the second level's butterflies (dft4_tw): 154 instead of
// 164 operations.  -DFMB_PLAIN_BUTTERFLIES restores the multiply-then-butterfly form of round 1.
// second level of dft16 (the first level - four dft4 over v[r0], v[4 + r0], v[8 + r0], v[12 + r0] - is done)
template <typename C> FMB_HD void dft16_level2(C *v) {
    typedef typename real_of<C>::type S;
    const S h = (S)0.70710678118654752440084436210485;
    const S c1 = (S)0.92387953251128675612818318939679;   // cos(pi/8)
    const S s1 = (S)0.38268343236508977172845998403040;   // sin(pi/8)
    dft4(v[0], v[1], v[2], v[3]);
    dft4_tw(v[4], v[5], v[6], v[7], c1, -s1, h, -h, s1, -c1);                       // k0 = 1: W16^1, W16^2, W16^3
    dft4_tw_mi(v[8], v[9], v[10], v[11], h, -h, -h, -h);                            // k0 = 2: W16^2, W16^4 = -i, W16^6
    dft4_tw(v[12], v[13], v[14], v[15], s1, -c1, -h, -h, -c1, s1);                  // k0 = 3: W16^3, W16^6, W16^9
}
// radix-4 butterfly whose FOUR inputs carry run-time twiddles w0 .. w3 (a stage twiddle, a spectrum value, ...):
// 28 operations instead of 16 (four complex multiplications) + 16
template <typename C> FMB_HD void dft4_tw4(C &v0, C &v1, C &v2, C &v3, C w0, C w1, C w2, C w3) {
    C t0, t1, t2, d;
    bfly_ww(v0, v2, w0.x, w0.y, w2.x, w2.y, t0, t1);
    bfly_ww(v1, v3, w1.x, w1.y, w3.x, w3.y, t2, d);
    v0 = cadd(t0, t2); v2 = csub(t0, t2);
    v1 = mk<C>(t1.x + d.y, t1.y - d.x);
    v3 = mk<C>(t1.x - d.y, t1.y + d.x);
}
template <typename C> FMB_HD void dft16(C *v) {
    typedef typename real_of<C>::type S;
    const S h = (S)0.70710678118654752440084436210485;
    const S c1 = (S)0.92387953251128675612818318939679;   // cos(pi/8)
    const S s1 = (S)0.38268343236508977172845998403040;   // sin(pi/8)
#pragma unroll
    for (int r0 = 0; r0 < 4; ++r0) dft4(v[r0], v[4 + r0], v[8 + r0], v[12 + r0]);   // y[r0][k0] at v[4*k0 + r0]
#ifndef FMB_PLAIN_BUTTERFLIES
    dft16_level2(v);
    (void)h; (void)c1; (void)s1;
#else
    // twiddles W16^{r0*k0} on v[4*k0 + r0]
    C a;
    // k0 = 1: W^1, W^2, W^3
    a = v[5];  v[5]  = mk<C>(a.x * c1 + a.y * s1, a.y * c1 - a.x * s1);               // W16^1 = c1 - i s1
    a = v[6];  v[6]  = mk<C>((a.x + a.y) * h, (a.y - a.x) * h);                       // W16^2 = (1 - i)/sqrt2
    a = v[7];  v[7]  = mk<C>(a.x * s1 + a.y * c1, a.y * s1 - a.x * c1);               // W16^3 = s1 - i c1
    // k0 = 2: W^2, W^4, W^6
    a = v[9];  v[9]  = mk<C>((a.x + a.y) * h, (a.y - a.x) * h);
    v[10] = cmul_mi(v[10]);                                                           // W16^4 = -i
    a = v[11]; v[11] = mk<C>((a.y - a.x) * h, -(a.x + a.y) * h);                      // W16^6 = (-1 - i)/sqrt2
    // k0 = 3: W^3, W^6, W^9
    a = v[13]; v[13] = mk<C>(a.x * s1 + a.y * c1, a.y * s1 - a.x * c1);               // W16^3
    a = v[14]; v[14] = mk<C>((a.y - a.x) * h, -(a.x + a.y) * h);                      // W16^6
    a = v[15]; v[15] = mk<C>(-a.x * c1 - a.y * s1, a.x * s1 - a.y * c1);              // W16^9 = -c1 + i s1
#pragma unroll
    for (int k0 = 0; k0 < 4; ++k0) dft4(v[4 * k0], v[4 * k0 + 1], v[4 * k0 + 2], v[4 * k0 + 3]);
#endif
}

// generic odd prime radix, constants taken from the unit-root table of the transform (wp[r] = e^{-2 pi i r/P})
template <typename C, int P> FMB_HD void dft_odd(C *v, const C *wp /* wp[1..(P-1)/2] valid */) {
    typedef typename real_of<C>::type S;
    constexpr int H = (P - 1) / 2;
    C a[H + 1], b[H + 1];
#pragma unroll
    for (int r = 1; r <= H; ++r) { a[r] = cadd(v[r], v[P - r]); b[r] = csub(v[r], v[P - r]); }
    C x0 = v[0];
    C sum = x0;
#pragma unroll
    for (int r = 1; r <= H; ++r) sum = cadd(sum, a[r]);
    v[0] = sum;
#pragma unroll
    for (int q = 1; q <= H; ++q) {
        S ar = x0.x, ai = x0.y, br = 0, bi = 0;
#pragma unroll
        for (int r = 1; r <= H; ++r) {
            int e = (r * q) % P;
            // W^{e} = cos - i sin; use symmetry W^{P-e} = conj(W^{e})
            S c = (e <= H) ? wp[e].x : wp[P - e].x;
            S s = (e <= H) ? -wp[e].y : wp[P - e].y;       // s = sin(2 pi e / P)
            ar += c * a[r].x; ai += c * a[r].y;
            br += s * b[r].y; bi -= s * b[r].x;             // (-i s) * b = (s b.y, -s b.x)
        }
        v[q] = mk<C>(ar + br, ai + bi);
        v[P - q] = mk<C>(ar - br, ai - bi);
    }
}

}  // namespace fmb
