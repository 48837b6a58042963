// Bit-faithful restatement of glibc 2.39 sinf / cosf for |x| < 120 (the only range the Costas loop of
// the reference, signal_functions.pyx:301, can produce: its phase is wrapped to +-2*pi every sample).
//
// Provenance of every constant and of the operation order: NOT glibc source, but the machine code and
// .rodata of THIS image's /usr/lib/x86_64-linux-gnu/libm.so.6 (glibc 2.39-0ubuntu8.5, build-id
// 0d9969fe206760d250ec30a5a9be18aefbf84ea8), read with objdump/readelf in round 1:
//   * sinf / cosf are IFUNCs; on CPUs with FMA+AVX2 (this container's Xeon and the B200 host) they resolve
//     to the FMA variants at 0x7e800 / 0x7e330, whose double-precision polynomial steps are contracted
//     into vfmadd exactly as written below (fma() here == one IEEE fused operation, as on the GPU);
//   * the 14-double table __sincosf_table[2] sits at 0xb8120 (signs, 2/pi*2^24, pi/2, c0,c1,s1,c2,s2,c3,s3,c4);
//     entry [1] negates the cosine polynomial.
// The non-FMA (SSE2) variant differs only in the last bit of the double intermediates (observable in the
// float result with probability ~2^-29 per call).  tests/test_sincosf_restatement.py pins this header
// against libm bit-for-bit on the CPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define URH_SC_HD __device__ __forceinline__
#define URH_DMUL(a, b) __dmul_rn((a), (b))
#define URH_DFMA(a, b, c) __fma_rn((a), (b), (c))
#define URH_D2I_RZ(x) __double2int_rz(x)
#define URH_D2F(x) __double2float_rn(x)
#define URH_SC_F2U(x) __float_as_uint(x)
#else
#if defined(__CUDACC__)
#define URH_SC_HD __host__ __device__ inline
#else
#define URH_SC_HD static inline
#endif
#define URH_DMUL(a, b) ((a) * (b))
#define URH_DFMA(a, b, c) fma((a), (b), (c))
#define URH_D2I_RZ(x) ((int32_t)(x))
#define URH_D2F(x) ((float)(x))
static inline uint32_t urh_sc_f2u(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
#define URH_SC_F2U(x) urh_sc_f2u(x)
#endif

#define URH_SC_HPI_INV 0x1.45f306dc9c883p+23
#define URH_SC_HPI 0x1.921fb54442d18p+0
#define URH_SC_C1 -0x1.ffffffd0c621cp-2
#define URH_SC_C2 0x1.55553e1068f19p-5
#define URH_SC_C3 -0x1.6c087e89a359dp-10
#define URH_SC_C4 0x1.99343027bf8c3p-16
#define URH_SC_S1 -0x1.555545995a603p-3
#define URH_SC_S2 0x1.1107605230bc4p-7
#define URH_SC_S3 -0x1.994eb3774cf24p-13

// sine polynomial on (x*sign, x^2): s = fma(x3, S1, x); result = fma(fma(S3, x2, S2), x3*x2, s)
URH_SC_HD float urh_sc_sin_poly(double xs, double x2) {
    const double s1p = URH_DFMA(URH_SC_S3, x2, URH_SC_S2);
    const double x3 = URH_DMUL(x2, xs);
    const double x5 = URH_DMUL(x2, x3);
    const double s = URH_DFMA(x3, URH_SC_S1, xs);
    return URH_D2F(URH_DFMA(s1p, x5, s));
}
// cosine polynomial; neg = use table[1] (all cosine coefficients negated)
URH_SC_HD float urh_sc_cos_poly(double x2, int neg) {
    const double sg = neg ? -1.0 : 1.0;
    const double x4 = URH_DMUL(x2, x2);
    const double c1p = URH_DFMA(sg * URH_SC_C1, x2, sg);
    const double c2p = URH_DFMA(sg * URH_SC_C4, x2, sg * URH_SC_C3);
    const double x6 = URH_DMUL(x2, x4);
    const double c = URH_DFMA(x4, sg * URH_SC_C2, c1p);
    return URH_D2F(URH_DFMA(c2p, x6, c));
}

// *ok = 0 when |y| >= 120 or non-finite (outside the restated range)
URH_SC_HD void urh_glibc_sincosf(float y, float* sn, float* cs, int* ok) {
    const uint32_t top = (URH_SC_F2U(y) >> 20) & 0x7ff;
    const double x = (double)y;
    *ok = 1;
    if (top <= 0x3f3) {  // |y| < pi/4
        if (top <= 0x397) {  // |y| < 2^-12
            *sn = y;
            *cs = 1.0f;
            return;
        }
        const double x2 = URH_DMUL(x, x);
        *sn = urh_sc_sin_poly(x, x2);
        *cs = urh_sc_cos_poly(x2, 0);
        return;
    }
    if (top > 0x42e) {  // |y| >= 120: reduce_large / inf / nan — not restated
        *ok = 0;
        *sn = 0.0f;
        *cs = 0.0f;
        return;
    }
    const double r = URH_DMUL(x, URH_SC_HPI_INV);
    const int32_t n = (URH_D2I_RZ(r) + 0x800000) >> 24;
    const double xr = URH_DFMA(-(double)n, URH_SC_HPI, x);  // vfnmadd: x - n*hpi, one rounding
    const double x2 = URH_DMUL(xr, xr);
    const int neg = (n & 2) ? 1 : 0;
    const double sign = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    const double xs = URH_DMUL(xr, sign);
    // sinf uses polynomial n, cosf uses polynomial n^1 (odd -> cosine polynomial)
    if (n & 1) {
        *sn = urh_sc_cos_poly(x2, neg);
        *cs = urh_sc_sin_poly(xs, x2);
    } else {
        *sn = urh_sc_sin_poly(xs, x2);
        *cs = urh_sc_cos_poly(x2, neg);
    }
}
