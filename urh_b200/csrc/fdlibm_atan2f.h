// Bit-faithful float32 restatement of glibc 2.39 atan2f / atanf (the fdlibm float algorithm,
// sysdeps/ieee754/flt-32/{e_atan2f.c,s_atanf.c}).
//
// Why: the reference's FSK demodulator (src/urh/cythonext/signal_functions.pyx:375-376) calls
// `atan2(tmp.imag, tmp.real)` on float operands in a C++ translation unit, which resolves to
// glibc's atan2f.  For the demodulated samples to be bit-identical to the reference, the device
// code has to execute exactly the same float32 operation sequence: every operation individually
// rounded (no FMA contraction), the same constants, the same branch thresholds.
//
// The header compiles for both device (intrinsics, never contracted) and host (plain operators;
// build with -ffp-contract=off) so that the CPU-only test-suite can check the restatement
// against libm's atan2f without a GPU (tests/test_atan2f_restatement.py).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define URH_HD __device__ __forceinline__
#define URH_MUL(a, b) __fmul_rn((a), (b))
#define URH_ADD(a, b) __fadd_rn((a), (b))
#define URH_SUB(a, b) __fsub_rn((a), (b))
#define URH_DIV(a, b) __fdiv_rn((a), (b))
#define URH_F2I(x) __float_as_int(x)
#define URH_I2F(x) __int_as_float(x)
#elif defined(__CUDACC__)
#define URH_HD __host__ __device__ inline
#define URH_MUL(a, b) ((a) * (b))
#define URH_ADD(a, b) ((a) + (b))
#define URH_SUB(a, b) ((a) - (b))
#define URH_DIV(a, b) ((a) / (b))
static inline int32_t urh_f2i_host(float x) { int32_t i; memcpy(&i, &x, 4); return i; }
static inline float urh_i2f_host(int32_t i) { float x; memcpy(&x, &i, 4); return x; }
#define URH_F2I(x) urh_f2i_host(x)
#define URH_I2F(x) urh_i2f_host(x)
#else
#define URH_HD static inline
#define URH_MUL(a, b) ((a) * (b))
#define URH_ADD(a, b) ((a) + (b))
#define URH_SUB(a, b) ((a) - (b))
#define URH_DIV(a, b) ((a) / (b))
static inline int32_t urh_f2i_host(float x) { int32_t i; memcpy(&i, &x, 4); return i; }
static inline float urh_i2f_host(int32_t i) { float x; memcpy(&x, &i, 4); return x; }
#define URH_F2I(x) urh_f2i_host(x)
#define URH_I2F(x) urh_i2f_host(x)
#endif

// Constants are the decimal literals of the fdlibm sources rounded to float32.
#define URH_ATANHI0 4.6364760399e-01f
#define URH_ATANHI1 7.8539812565e-01f
#define URH_ATANHI2 9.8279368877e-01f
#define URH_ATANHI3 1.5707962513e+00f
#define URH_ATANLO0 5.0121582440e-09f
#define URH_ATANLO1 3.7748947079e-08f
#define URH_ATANLO2 3.4473217170e-08f
#define URH_ATANLO3 7.5497894159e-08f
#define URH_AT0 3.3333334327e-01f
#define URH_AT1 -2.0000000298e-01f
#define URH_AT2 1.4285714924e-01f
#define URH_AT3 -1.1111110449e-01f
#define URH_AT4 9.0908870101e-02f
#define URH_AT5 -7.6918758452e-02f
#define URH_AT6 6.6610731184e-02f
#define URH_AT7 -5.8335702866e-02f
#define URH_AT8 4.9768779427e-02f
#define URH_AT9 -3.6531571299e-02f
#define URH_AT10 1.6285819933e-02f
#define URH_PI 3.1415927410e+00f
#define URH_PI_O_2 1.5707963705e+00f
#define URH_PI_O_4 7.8539818525e-01f
#define URH_PI_LO -8.7422776573e-08f
#define URH_TINY 1.0e-30f

// atanf for a NON-NEGATIVE, finite-or-inf, non-NaN argument (all atan2f needs: it passes |y/x|).
// ax >= 0.  Returns atanf(ax) exactly as glibc 2.39 computes it.
URH_HD float urh_atanf_pos(float ax) {
    const int32_t ix = URH_F2I(ax);
    if (ix >= 0x4c000000) {  // |x| >= 2^25 (inf included; NaN excluded by the caller)
        return URH_ADD(URH_ATANHI3, URH_ATANLO3);
    }
    float x, hi, lo;
    if (ix < 0x3ee00000) {         // |x| < 0.4375 : no argument reduction
        if (ix < 0x31000000) return ax;  // |x| < 2^-29
        const float z = URH_MUL(ax, ax);
        const float w = URH_MUL(z, z);
        const float s1 = URH_MUL(z, URH_ADD(URH_AT0, URH_MUL(w, URH_ADD(URH_AT2, URH_MUL(w, URH_ADD(URH_AT4,
                          URH_MUL(w, URH_ADD(URH_AT6, URH_MUL(w, URH_ADD(URH_AT8, URH_MUL(w, URH_AT10)))))))))));
        const float s2 = URH_MUL(w, URH_ADD(URH_AT1, URH_MUL(w, URH_ADD(URH_AT3, URH_MUL(w, URH_ADD(URH_AT5,
                          URH_MUL(w, URH_ADD(URH_AT7, URH_MUL(w, URH_AT9)))))))));
        return URH_SUB(ax, URH_MUL(ax, URH_ADD(s1, s2)));
    }
    // One division with interval-dependent numerator / denominator.
    float num, den;
    if (ix < 0x3f980000) {         // |x| < 1.1875
        if (ix < 0x3f300000) {     // 7/16 <= |x| < 11/16
            num = URH_SUB(URH_MUL(2.0f, ax), 1.0f); den = URH_ADD(2.0f, ax);
            hi = URH_ATANHI0; lo = URH_ATANLO0;
        } else {                   // 11/16 <= |x| < 19/16
            num = URH_SUB(ax, 1.0f); den = URH_ADD(ax, 1.0f);
            hi = URH_ATANHI1; lo = URH_ATANLO1;
        }
    } else {
        if (ix < 0x401c0000) {     // |x| < 2.4375
            num = URH_SUB(ax, 1.5f); den = URH_ADD(1.0f, URH_MUL(1.5f, ax));
            hi = URH_ATANHI2; lo = URH_ATANLO2;
        } else {                   // 2.4375 <= |x| < 2^25
            num = -1.0f; den = ax;
            hi = URH_ATANHI3; lo = URH_ATANLO3;
        }
    }
    x = URH_DIV(num, den);
    const float z = URH_MUL(x, x);
    const float w = URH_MUL(z, z);
    const float s1 = URH_MUL(z, URH_ADD(URH_AT0, URH_MUL(w, URH_ADD(URH_AT2, URH_MUL(w, URH_ADD(URH_AT4,
                      URH_MUL(w, URH_ADD(URH_AT6, URH_MUL(w, URH_ADD(URH_AT8, URH_MUL(w, URH_AT10)))))))))));
    const float s2 = URH_MUL(w, URH_ADD(URH_AT1, URH_MUL(w, URH_ADD(URH_AT3, URH_MUL(w, URH_ADD(URH_AT5,
                      URH_MUL(w, URH_ADD(URH_AT7, URH_MUL(w, URH_AT9)))))))));
    return URH_SUB(hi, URH_SUB(URH_SUB(URH_MUL(x, URH_ADD(s1, s2)), lo), x));
}

// Full-sign atanf (used when atan2f is called with x == 1.0f).
URH_HD float urh_atanf(float x) {
    const int32_t hx = URH_F2I(x);
    const int32_t ix = hx & 0x7fffffff;
    if (ix > 0x7f800000) return URH_ADD(x, x);  // NaN
    const float r = urh_atanf_pos(URH_I2F(ix));
    // glibc: small/identity branches return x itself (sign kept); reduced branches negate z.
    return (hx < 0) ? -r : r;
}

URH_HD float urh_atan2f(float y, float x) {
    const int32_t hx = URH_F2I(x), hy = URH_F2I(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return URH_ADD(x, y);  // NaN
    if (hx == 0x3f800000) return urh_atanf(y);                      // x == 1.0
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);              // 2*sign(x) + sign(y)
    if (iy == 0) {                                                   // y == +-0
        if (m < 2) return y;
        return (m == 2) ? URH_ADD(URH_PI, URH_TINY) : URH_SUB(-URH_PI, URH_TINY);
    }
    if (ix == 0) return (hy < 0) ? URH_SUB(-URH_PI_O_2, URH_TINY) : URH_ADD(URH_PI_O_2, URH_TINY);
    if (ix == 0x7f800000) {                                          // x == +-inf
        if (iy == 0x7f800000) {
            switch (m) {
                case 0: return URH_ADD(URH_PI_O_4, URH_TINY);
                case 1: return URH_SUB(-URH_PI_O_4, URH_TINY);
                case 2: return URH_ADD(URH_MUL(3.0f, URH_PI_O_4), URH_TINY);
                default: return URH_SUB(URH_MUL(-3.0f, URH_PI_O_4), URH_TINY);
            }
        } else {
            switch (m) {
                case 0: return 0.0f;
                case 1: return -0.0f;
                case 2: return URH_ADD(URH_PI, URH_TINY);
                default: return URH_SUB(-URH_PI, URH_TINY);
            }
        }
    }
    if (iy == 0x7f800000) return (hy < 0) ? URH_SUB(-URH_PI_O_2, URH_TINY) : URH_ADD(URH_PI_O_2, URH_TINY);
    const int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = URH_ADD(URH_PI_O_2, URH_MUL(0.5f, URH_PI_LO));  // |y/x| > 2^60
    else if (hx < 0 && k < -60) z = 0.0f;                            // |y|/x < -2^60
    else {
        const float q = URH_DIV(y, x);
        z = urh_atanf_pos(URH_I2F(URH_F2I(q) & 0x7fffffff));
    }
    switch (m) {
        case 0: return z;
        case 1: return URH_I2F(URH_F2I(z) ^ (int32_t)0x80000000);
        case 2: return URH_SUB(URH_PI, URH_SUB(z, URH_PI_LO));
        default: return URH_SUB(URH_SUB(z, URH_PI_LO), URH_PI);
    }
}

// ---- branch-light formulation (same results, fewer instructions on the GPU) -------------------------
// urh_atan2f_v2(y, x) == urh_atan2f(y, x) bit-for-bit (checked on the CPU by
// tests/test_atan2f_restatement.py against libm).  Differences in structure only:
//  * the zero / one / exponent-gap special cases of e_atan2f.c are folded into the general path where
//    the general path provably yields the same float (y == +-0 with x != 0, x == 1, |k| > 60);
//    only x == +-0 and non-finite operands take the literal reference path;
//  * |y/x| < 0.4375 (every narrow-band FSK sample) needs no second division and no table;
//  * the four reduction intervals share one num/den/poly evaluation through selected constants:
//    num = fl(fl(a*q) + b), den = fl(fl(c*q) + d) reproduces (2q-1)/(2+q), (q-1)/(q+1),
//    (q-1.5)/(1+1.5q) and -1/q with the reference's roundings (a*q is exact for a in {0,1,2}).
URH_HD float urh_atan2f_v2(float y, float x) {
    const int32_t hx = URH_F2I(x), hy = URH_F2I(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (((uint32_t)(ix - 1) >= 0x7f7fffffu) | (iy >= 0x7f800000)) return urh_atan2f(y, x);
    const float q = URH_DIV(URH_I2F(iy), URH_I2F(ix));
    const int32_t iq = URH_F2I(q);
    float z;
    if (iq < 0x3ee00000) {
        const float z2 = URH_MUL(q, q);
        const float w = URH_MUL(z2, z2);
        const float s1 = URH_MUL(z2, URH_ADD(URH_AT0, URH_MUL(w, URH_ADD(URH_AT2, URH_MUL(w, URH_ADD(URH_AT4,
                          URH_MUL(w, URH_ADD(URH_AT6, URH_MUL(w, URH_ADD(URH_AT8, URH_MUL(w, URH_AT10)))))))))));
        const float s2 = URH_MUL(w, URH_ADD(URH_AT1, URH_MUL(w, URH_ADD(URH_AT3, URH_MUL(w, URH_ADD(URH_AT5,
                          URH_MUL(w, URH_ADD(URH_AT7, URH_MUL(w, URH_AT9)))))))));
        z = URH_SUB(q, URH_MUL(q, URH_ADD(s1, s2)));
    } else if (iq >= 0x4c000000) {
        z = URH_ADD(URH_ATANHI3, URH_ATANLO3);
    } else {
        float a, b, c, d, hi, lo;
        if (iq < 0x3f980000) {
            if (iq < 0x3f300000) { a = 2.0f; b = -1.0f; c = 1.0f; d = 2.0f; hi = URH_ATANHI0; lo = URH_ATANLO0; }
            else { a = 1.0f; b = -1.0f; c = 1.0f; d = 1.0f; hi = URH_ATANHI1; lo = URH_ATANLO1; }
        } else {
            if (iq < 0x401c0000) { a = 1.0f; b = -1.5f; c = 1.5f; d = 1.0f; hi = URH_ATANHI2; lo = URH_ATANLO2; }
            else { a = 0.0f; b = -1.0f; c = 1.0f; d = 0.0f; hi = URH_ATANHI3; lo = URH_ATANLO3; }
        }
        const float num = URH_ADD(URH_MUL(a, q), b);
        const float den = URH_ADD(URH_MUL(c, q), d);
        const float xr = URH_DIV(num, den);
        const float z2 = URH_MUL(xr, xr);
        const float w = URH_MUL(z2, z2);
        const float s1 = URH_MUL(z2, URH_ADD(URH_AT0, URH_MUL(w, URH_ADD(URH_AT2, URH_MUL(w, URH_ADD(URH_AT4,
                          URH_MUL(w, URH_ADD(URH_AT6, URH_MUL(w, URH_ADD(URH_AT8, URH_MUL(w, URH_AT10)))))))))));
        const float s2 = URH_MUL(w, URH_ADD(URH_AT1, URH_MUL(w, URH_ADD(URH_AT3, URH_MUL(w, URH_ADD(URH_AT5,
                          URH_MUL(w, URH_ADD(URH_AT7, URH_MUL(w, URH_AT9)))))))));
        z = URH_SUB(hi, URH_SUB(URH_SUB(URH_MUL(xr, URH_ADD(s1, s2)), lo), xr));
    }
    float r = z;
    if (hx < 0) r = URH_SUB(URH_PI, URH_SUB(z, URH_PI_LO));
    return (hy < 0) ? URH_I2F(URH_F2I(r) ^ (int32_t)0x80000000) : r;
}
