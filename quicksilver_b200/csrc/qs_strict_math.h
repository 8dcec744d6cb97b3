/* qs_strict_math.h -- portable log / sin / cos used by the VALIDATION build.
 *
 * The reference calls libm's log (src/MC_Segment_Outcome.cc:83, src/CollisionEvent.cc:44,
 * src/MC_SourceNow.cc:112) and sin/cos (src/CollisionEvent.cc:31-32) inside the tracking
 * loop.  glibc and CUDA's libdevice round these differently in the last bit, so a
 * device history can never be compared bit-for-bit with a host history through them.
 * This header restates the three functions with nothing but IEEE-754 double + - * / and
 * integer bit operations (classic Cody-Waite reduction + minimax polynomials), so the
 * SAME source compiled by gcc (-ffp-contract=off) and by nvcc (--fmad=false) produces
 * the same bits on the host and on sm_100a.  Accuracy is ~1 ulp, which is all the
 * physics needs.  The FAST build uses the same functions, contracted to FMAs by nvcc
 * (track_physics.cuh: m_log / m_sincos): the CUDA math library's out-of-line argument
 * reduction is never needed for the arguments the tracking loop passes.
 * Algorithms and coefficients: the classic public-domain fdlibm lineage (Sun Microsystems'
 * e_log.c, k_sin.c, k_cos.c: Cody-Waite reduction by ln2 / pi/2 in two pieces, minimax
 * polynomials L1..L7, S1..S6, C1..C6), restated for a restricted argument range.
 */
#ifndef QS_STRICT_MATH_H
#define QS_STRICT_MATH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define QS_HD __host__ __device__ __forceinline__
#else
#define QS_HD static inline
#endif

QS_HD uint64_t qs_f64_bits(double x)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}

QS_HD double qs_bits_f64(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}

/* natural logarithm, x > 0 finite (x == 0 returns -inf, x < 0 returns NaN) */
QS_HD double qs_strict_log(double x)
{
    const double ln2_hi = 6.93147180369123816490e-01;
    const double ln2_lo = 1.90821492927058770002e-10;
    const double L1 = 6.666666666666735130e-01, L2 = 3.999999999940941908e-01,
                 L3 = 2.857142874366239149e-01, L4 = 2.222219843214978396e-01,
                 L5 = 1.818357216161805012e-01, L6 = 1.531383769920937332e-01,
                 L7 = 1.479819860511658591e-01;
    uint64_t ux = qs_f64_bits(x);
    int k = 0;
    if ((ux >> 63) != 0 || (ux << 1) == 0)
    {
        if ((ux << 1) == 0) return qs_bits_f64(0xFFF0000000000000ull);      /* -inf */
        return qs_bits_f64(0x7FF8000000000000ull);                          /* NaN  */
    }
    if ((ux >> 52) == 0)                       /* subnormal: scale up by 2^54 */
    {
        x = x * 18014398509481984.0;
        ux = qs_f64_bits(x);
        k = -54;
    }
    if ((ux >> 52) == 0x7FF) return x;         /* +inf / NaN */
    /* x = 2^k * m,  m in [sqrt(2)/2, sqrt(2)) */
    uint32_t hx = (uint32_t)(ux >> 32);
    hx += 0x3FF00000u - 0x3FE6A09Eu;
    k += (int)(hx >> 20) - 0x3FF;
    hx = (hx & 0x000FFFFFu) + 0x3FE6A09Eu;
    double m = qs_bits_f64(((uint64_t)hx << 32) | (ux & 0xFFFFFFFFull));
    double f = m - 1.0;
    double hfsq = 0.5 * f * f;
    double s = f / (2.0 + f);
    double z = s * s;
    double w = z * z;
    double t1 = w * (L2 + w * (L4 + w * L6));
    double t2 = z * (L1 + w * (L3 + w * (L5 + w * L7)));
    double R = t2 + t1;
    double dk = (double)k;
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

/* sin and cos of phi for 0 <= phi < ~8 (the tracking loop only ever passes
 * phi = 2 * 3.14159265 * r with r in [0,1), src/CollisionEvent.cc:30). */
QS_HD void qs_strict_sincos(double phi, double* sn, double* cs)
{
    const double two_over_pi = 6.36619772367581382433e-01;
    const double pio2_hi = 1.57079632673412561417e+00;   /* first 33 bits of pi/2 */
    const double pio2_lo = 6.07710050650619224932e-11;   /* pi/2 - pio2_hi        */
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    int n = (int)(phi * two_over_pi + 0.5);
    double dn = (double)n;
    double y = (phi - dn * pio2_hi) - dn * pio2_lo;      /* |y| <= pi/4 (+ a hair) */
    double z = y * y;
    double ps = S1 + z * (S2 + z * (S3 + z * (S4 + z * (S5 + z * S6))));
    double pc = C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6))));
    double s = y + (y * z) * ps;
    double c = (1.0 - 0.5 * z) + (z * z) * pc;
    switch (n & 3)
    {
        case 0:  *sn =  s; *cs =  c; break;
        case 1:  *sn =  c; *cs = -s; break;
        case 2:  *sn = -s; *cs = -c; break;
        default: *sn = -c; *cs =  s; break;
    }
}

#endif /* QS_STRICT_MATH_H */
