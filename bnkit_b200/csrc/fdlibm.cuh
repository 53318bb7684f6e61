// fdlibm.cuh -- fdlibm's e_exp.c / e_log.c on the device (the algorithm java.lang.StrictMath is specified to be).
// IEEE-754 arithmetic only: in a unit compiled with -fmad=false the results are bit-identical to the oracle's
// om_exp / om_log (the C restatement under oracle/).  Included by tables.cu and indel_tables.cu.
#pragma once
#include "device_common.cuh"

namespace bnk {

__device__ __forceinline__ double with_hi(double x, int hi) { return __hiloint2double(hi, __double2loint(x)); }

static __device__ double fd_exp(double x)
{
    const double huge = 1.0e+300, twom1000 = 9.33263618503218878990e-302,
                 o_threshold = 7.09782712893383973096e+02, u_threshold = -7.45133219101941108420e+02,
                 ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
                 invln2 = 1.44269504088896338700e+00, P1 = 1.66666666666666019037e-01,
                 P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    double y, hi = 0.0, lo = 0.0, c, t;
    int k = 0;
    unsigned hx = (unsigned)__double2hiint(x);
    const int xsb = (int)((hx >> 31) & 1u);
    hx &= 0x7fffffffu;
    if (hx >= 0x40862E42u) {
        if (hx >= 0x7ff00000u) {
            if (((hx & 0xfffffu) | (unsigned)__double2loint(x)) != 0u) return x + x;
            return xsb == 0 ? x : 0.0;
        }
        if (x > o_threshold) return huge * huge;
        if (x < u_threshold) return twom1000 * twom1000;
    }
    if (hx > 0x3fd62e42u) {
        if (hx < 0x3FF0A2B2u) {
            hi = xsb ? x + ln2HI : x - ln2HI;
            lo = xsb ? -ln2LO : ln2LO;
            k = 1 - xsb - xsb;
        } else {
            k = (int)(invln2 * x + (xsb ? -0.5 : 0.5));
            t = (double)k;
            hi = x - t * ln2HI;
            lo = t * ln2LO;
        }
        x = hi - lo;
    } else if (hx < 0x3e300000u) {
        if (huge + x > 1.0) return 1.0 + x;
    }
    t = x * x;
    c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
    y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
    if (k >= -1021) return with_hi(y, __double2hiint(y) + (k << 20));
    y = with_hi(y, __double2hiint(y) + ((k + 1000) << 20));
    return y * twom1000;
}

static __device__ double fd_log(double x)
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 two54 = 1.80143985094819840000e+16, Lg1 = 6.666666666666735130e-01,
                 Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01,
                 Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
    double hfsq, f, s, z, R, w, t1, t2, dk;
    int k = 0, i, j;
    int hx = __double2hiint(x);
    const unsigned lx = (unsigned)__double2loint(x);
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | (int)lx) == 0) return -CUDART_INF;
        if (hx < 0) return CUDART_NAN;
        k -= 54;
        x *= two54;
        hx = __double2hiint(x);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    x = with_hi(x, hx | (i ^ 0x3ff00000));
    k += (i >> 20);
    f = x - 1.0;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) {
            if (k == 0) return 0.0;
            dk = (double)k;
            return dk * ln2_hi + dk * ln2_lo;
        }
        R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        dk = (double)k;
        return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    s = f / (2.0 + f);
    dk = (double)k;
    z = s * s;
    i = hx - 0x6147a;
    w = z * z;
    j = 0x6b851 - hx;
    t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    R = t2 + t1;
    if (i > 0) {
        hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

}  // namespace bnk
