/*
 * bnk_oracle.c -- CPU ORACLE, part 1: elementary functions, substitution
 * models, the real general eigen-solver and P(t).  Test infrastructure only
 * (see bnk_oracle.h).  Compile with -ffp-contract=off: every multiply and add
 * below must round separately, as the JVM does.
 *
 * Follows (reference file:line):
 *   src/bn/ctmc/SubstModel.java:80-110   constructor: R from S*F (upper triangle) or Q as given
 *   src/bn/ctmc/SubstModel.java:192-202  makeValid
 *   src/bn/ctmc/SubstModel.java:209-219  normalize
 *   src/bn/ctmc/SubstModel.java:303-327  getProbs
 *   src/bn/math/Matrix.java:42-60        Exp: elmhes -> eltran -> hqr2 -> luinverse
 *   src/bn/math/Matrix.java:170-215      elmhes
 *   src/bn/math/Matrix.java:232-258      eltran
 *   src/bn/math/Matrix.java:268-280      mcdiv
 *   src/bn/math/Matrix.java:311-722      hqr2
 *   src/bn/math/Matrix.java:733-837      luinverse
 *   src/bn/ctmc/matrix/{LG,JTT,WAG,Dayhoff,JC,Yang}.java  tables (via model_tables.h)
 *   src/bn/ctmc/SubstNode.java:100-113 + src/bn/prob/EnumDistrib.java:66-78,363-371  root prior F/sum(F)
 *   src/dat/phylo/IdxTree.java:411-464   getPrunedIndex (orphan removal)
 *
 * exp/log: java.lang.Math.exp/log are specified to <= 1 ulp; java.lang.StrictMath
 * is specified to BE fdlibm.  The oracle restates fdlibm's e_exp.c / e_log.c
 * (Sun Microsystems 1993, freely distributable) so that results are reproducible
 * across machines; HotSpot's Math intrinsics may differ from it in the last bit.
 */
#include "bnk_oracle.h"
#include "model_tables.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ------------------------------------------------------------------ */
/* fdlibm exp / log                                                    */
/* ------------------------------------------------------------------ */
static inline uint64_t d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
static inline int32_t hi_word(double x) { return (int32_t)(d2u(x) >> 32); }
static inline uint32_t lo_word(double x) { return (uint32_t)d2u(x); }
static inline double with_hi(double x, int32_t hi) {
    return u2d(((uint64_t)(uint32_t)hi << 32) | (d2u(x) & 0xffffffffu));
}

double om_exp(double x)
{
    static const double halF[2] = {0.5, -0.5},
        huge = 1.0e+300, twom1000 = 9.33263618503218878990e-302,
        o_threshold = 7.09782712893383973096e+02,
        u_threshold = -7.45133219101941108420e+02,
        ln2HI[2] = {6.93147180369123816490e-01, -6.93147180369123816490e-01},
        ln2LO[2] = {1.90821492927058770002e-10, -1.90821492927058770002e-10},
        invln2 = 1.44269504088896338700e+00,
        P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
        P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
        P5 = 4.13813679705723846039e-08;
    double y, hi = 0.0, lo = 0.0, c, t;
    int32_t k = 0, xsb;
    uint32_t hx;

    hx = (uint32_t)hi_word(x);
    xsb = (int32_t)((hx >> 31) & 1);
    hx &= 0x7fffffff;
    if (hx >= 0x40862E42) {
        if (hx >= 0x7ff00000) {
            if (((hx & 0xfffff) | lo_word(x)) != 0) return x + x;
            return (xsb == 0) ? x : 0.0;
        }
        if (x > o_threshold) return huge * huge;
        if (x < u_threshold) return twom1000 * twom1000;
    }
    if (hx > 0x3fd62e42) {
        if (hx < 0x3FF0A2B2) {
            hi = x - ln2HI[xsb];
            lo = ln2LO[xsb];
            k = 1 - xsb - xsb;
        } else {
            k = (int32_t)(invln2 * x + halF[xsb]);
            t = k;
            hi = x - t * ln2HI[0];
            lo = t * ln2LO[0];
        }
        x = hi - lo;
    } else if (hx < 0x3e300000) {
        if (huge + x > 1.0) return 1.0 + x;
    } else
        k = 0;
    t = x * x;
    c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
    y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
    if (k >= -1021) {
        return with_hi(y, hi_word(y) + (k << 20));
    } else {
        y = with_hi(y, hi_word(y) + ((k + 1000) << 20));
        return y * twom1000;
    }
}

double om_log(double x)
{
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
        two54 = 1.80143985094819840000e+16,
        Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
        Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
        Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
        Lg7 = 1.479819860511658591e-01;
    static const double zero = 0.0;
    double hfsq, f, s, z, R, w, t1, t2, dk;
    int32_t k, hx, i, j;
    uint32_t lx;

    hx = hi_word(x);
    lx = lo_word(x);
    k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -two54 / zero;
        if (hx < 0) return (x - x) / zero;
        k -= 54;
        x *= two54;
        hx = hi_word(x);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    x = with_hi(x, hx | (i ^ 0x3ff00000));
    k += (i >> 20);
    f = x - 1.0;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == zero) {
            if (k == 0) return zero;
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
    } else {
        if (k == 0) return f - s * (f - R);
        return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
    }
}

/* Factorize.java:1199-1210; isLOG0 = Double.isInfinite (AbstractFactor.java:741-745) */
double om_logsum(double logx, double logy)
{
    int x0 = isinf(logx), y0 = isinf(logy);
    if (x0 && !y0) return logy;
    if (!x0 && y0) return logx;
    if (x0 && y0) return -INFINITY;
    if (logx > logy) return logx + om_log(1.0 + om_exp(logy - logx));
    return logy + om_log(1.0 + om_exp(logx - logy));
}

/* ------------------------------------------------------------------ */
/* eigen-solver (Matrix.java)                                          */
/* ------------------------------------------------------------------ */
#define A_(i, j) a[(i) * n + (j)]
#define H_(i, j) h[(i) * n + (j)]
#define Z_(i, j) zz[(i) * n + (j)]

/* Matrix.java:170-215 */
static void elmhes(double *a, int *ordr, int n)
{
    int m, j, i;
    double y, x;
    for (i = 0; i < n; i++) ordr[i] = 0;
    for (m = 2; m < n; m++) {
        x = 0.0;
        i = m;
        for (j = m; j <= n; j++) {
            if (fabs(A_(j - 1, m - 2)) > fabs(x)) {
                x = A_(j - 1, m - 2);
                i = j;
            }
        }
        ordr[m - 1] = i;
        if (i != m) {
            for (j = m - 2; j < n; j++) {
                y = A_(i - 1, j);
                A_(i - 1, j) = A_(m - 1, j);
                A_(m - 1, j) = y;
            }
            for (j = 0; j < n; j++) {
                y = A_(j, i - 1);
                A_(j, i - 1) = A_(j, m - 1);
                A_(j, m - 1) = y;
            }
        }
        if (x != 0.0) {
            for (i = m; i < n; i++) {
                y = A_(i, m - 2);
                if (y != 0.0) {
                    y /= x;
                    A_(i, m - 2) = y;
                    for (j = m - 1; j < n; j++) A_(i, j) -= y * A_(m - 1, j);
                    for (j = 0; j < n; j++) A_(j, m - 1) += y * A_(j, i);
                }
            }
        }
    }
}

/* Matrix.java:232-258 */
static void eltran(const double *a, double *zz, const int *ordr, int n)
{
    int i, j, m;
    for (i = 0; i < n; i++) {
        for (j = i + 1; j < n; j++) {
            Z_(i, j) = 0.0;
            Z_(j, i) = 0.0;
        }
        Z_(i, i) = 1.0;
    }
    if (n <= 2) return;
    for (m = n - 1; m >= 2; m--) {
        for (i = m; i < n; i++) Z_(i, m - 1) = A_(i, m - 2);
        i = ordr[m - 1];
        if (i != m) {
            for (j = m - 1; j < n; j++) {
                Z_(m - 1, j) = Z_(i - 1, j);
                Z_(i - 1, j) = 0.0;
            }
            Z_(i - 1, m - 1) = 1.0;
        }
    }
}

/* Matrix.java:268-280 */
static void mcdiv(double ar, double ai, double br, double bi, double *c)
{
    double s, ars, ais, brs, bis;
    s = fabs(br) + fabs(bi);
    ars = ar / s;
    ais = ai / s;
    brs = br / s;
    bis = bi / s;
    s = brs * brs + bis * bis;
    c[0] = (ars * brs + ais * bis) / s;
    c[1] = (ais * brs - ars * bis) / s;
}

/* Matrix.java:311-722.  Returns 0, or -1 when the iteration limit is hit
 * (the reference throws ArithmeticException there). */
static int hqr2(int n, int low, int hgh, double *h, double *zz, double *wr, double *wi)
{
    int i, j, k, l = 0, m, en, na, itn, its;
    double p = 0, q = 0, r = 0, s = 0, t, w, x = 0, y, ra, sa, vi, vr, z = 0, norm, tst1, tst2;
    double c[2];
    int notLast;

    norm = 0.0;
    k = 1;
    for (i = 0; i < n; i++) {
        for (j = k - 1; j < n; j++) norm += fabs(H_(i, j));
        k = i + 1;
        if (i + 1 < low || i + 1 > hgh) {
            wr[i] = H_(i, i);
            wi[i] = 0.0;
        }
    }
    en = hgh;
    t = 0.0;
    itn = n * 30;
    while (en >= low) {
        its = 0;
        na = en - 1;
        while (en >= 1) {
            int fullLoop = 1;
            for (l = en; l > low; l--) {
                s = fabs(H_(l - 2, l - 2)) + fabs(H_(l - 1, l - 1));
                if (s == 0.0) s = norm;
                tst1 = s;
                tst2 = tst1 + fabs(H_(l - 1, l - 2));
                if (tst2 == tst1) {
                    fullLoop = 0;
                    break;
                }
            }
            if (fullLoop) l = low;
            x = H_(en - 1, en - 1);
            if (l == en || l == na) break;
            if (itn == 0) return -1;
            y = H_(na - 1, na - 1);
            w = H_(en - 1, na - 1) * H_(na - 1, en - 1);
            if (its == 10 || its == 20) {
                t += x;
                for (i = low - 1; i < en; i++) H_(i, i) -= x;
                s = fabs(H_(en - 1, na - 1)) + fabs(H_(na - 1, en - 3));
                x = 0.75 * s;
                y = x;
                w = -0.4375 * s * s;
            }
            its++;
            itn--;
            for (m = en - 2; m >= l; m--) {
                z = H_(m - 1, m - 1);
                r = x - z;
                s = y - z;
                p = (r * s - w) / H_(m, m - 1) + H_(m - 1, m);
                q = H_(m, m) - z - r - s;
                r = H_(m + 1, m);
                s = fabs(p) + fabs(q) + fabs(r);
                p /= s;
                q /= s;
                r /= s;
                if (m == l) break;
                tst1 = fabs(p) * (fabs(H_(m - 2, m - 2)) + fabs(z) + fabs(H_(m, m)));
                tst2 = tst1 + fabs(H_(m - 1, m - 2)) * (fabs(q) + fabs(r));
                if (tst2 == tst1) break;
            }
            for (i = m + 2; i <= en; i++) {
                H_(i - 1, i - 3) = 0.0;
                if (i != m + 2) H_(i - 1, i - 4) = 0.0;
            }
            for (k = m; k <= na; k++) {
                notLast = (k != na);
                if (k != m) {
                    p = H_(k - 1, k - 2);
                    q = H_(k, k - 2);
                    r = 0.0;
                    if (notLast) r = H_(k + 1, k - 2);
                    x = fabs(p) + fabs(q) + fabs(r);
                    if (x != 0.0) {
                        p /= x;
                        q /= x;
                        r /= x;
                    }
                }
                if (x != 0.0) {
                    if (p < 0.0)
                        s = -sqrt(p * p + q * q + r * r);
                    else
                        s = sqrt(p * p + q * q + r * r);
                    if (k != m)
                        H_(k - 1, k - 2) = -s * x;
                    else if (l != m)
                        H_(k - 1, k - 2) = -H_(k - 1, k - 2);
                    p += s;
                    x = p / s;
                    y = q / s;
                    z = r / s;
                    q /= p;
                    r /= p;
                    if (!notLast) {
                        for (j = k - 1; j < n; j++) {
                            p = H_(k - 1, j) + q * H_(k, j);
                            H_(k - 1, j) -= p * x;
                            H_(k, j) -= p * y;
                        }
                        j = (en < (k + 3)) ? en : (k + 3);
                        for (i = 0; i < j; i++) {
                            p = x * H_(i, k - 1) + y * H_(i, k);
                            H_(i, k - 1) -= p;
                            H_(i, k) -= p * q;
                        }
                        for (i = low - 1; i < hgh; i++) {
                            p = x * Z_(i, k - 1) + y * Z_(i, k);
                            Z_(i, k - 1) -= p;
                            Z_(i, k) -= p * q;
                        }
                    } else {
                        for (j = k - 1; j < n; j++) {
                            p = H_(k - 1, j) + q * H_(k, j) + r * H_(k + 1, j);
                            H_(k - 1, j) -= p * x;
                            H_(k, j) -= p * y;
                            H_(k + 1, j) -= p * z;
                        }
                        j = (en < (k + 3)) ? en : (k + 3);
                        for (i = 0; i < j; i++) {
                            p = x * H_(i, k - 1) + y * H_(i, k) + z * H_(i, k + 1);
                            H_(i, k - 1) -= p;
                            H_(i, k) -= p * q;
                            H_(i, k + 1) -= p * r;
                        }
                        for (i = low - 1; i < hgh; i++) {
                            p = x * Z_(i, k - 1) + y * Z_(i, k) + z * Z_(i, k + 1);
                            Z_(i, k - 1) -= p;
                            Z_(i, k) -= p * q;
                            Z_(i, k + 1) -= p * r;
                        }
                    }
                }
            }
        }
        if (l == en) {
            H_(en - 1, en - 1) = x + t;
            wr[en - 1] = H_(en - 1, en - 1);
            wi[en - 1] = 0.0;
            en = na;
            continue;
        }
        y = H_(na - 1, na - 1);
        w = H_(en - 1, na - 1) * H_(na - 1, en - 1);
        p = (y - x) / 2.0;
        q = p * p + w;
        z = sqrt(fabs(q));
        H_(en - 1, en - 1) = x + t;
        x = H_(en - 1, en - 1);
        H_(na - 1, na - 1) = y + t;
        if (q >= 0.0) {
            if (p < 0.0)
                z = p - fabs(z);
            else
                z = p + fabs(z);
            wr[na - 1] = x + z;
            wr[en - 1] = wr[na - 1];
            if (z != 0.0) wr[en - 1] = x - w / z;
            wi[na - 1] = 0.0;
            wi[en - 1] = 0.0;
            x = H_(en - 1, na - 1);
            s = fabs(x) + fabs(z);
            p = x / s;
            q = z / s;
            r = sqrt(p * p + q * q);
            p /= r;
            q /= r;
            for (j = na - 1; j < n; j++) {
                z = H_(na - 1, j);
                H_(na - 1, j) = q * z + p * H_(en - 1, j);
                H_(en - 1, j) = q * H_(en - 1, j) - p * z;
            }
            for (i = 0; i < en; i++) {
                z = H_(i, na - 1);
                H_(i, na - 1) = q * z + p * H_(i, en - 1);
                H_(i, en - 1) = q * H_(i, en - 1) - p * z;
            }
            for (i = low - 1; i < hgh; i++) {
                z = Z_(i, na - 1);
                Z_(i, na - 1) = q * z + p * Z_(i, en - 1);
                Z_(i, en - 1) = q * Z_(i, en - 1) - p * z;
            }
        } else {
            wr[na - 1] = x + p;
            wr[en - 1] = x + p;
            wi[na - 1] = z;
            wi[en - 1] = -z;
        }
        en -= 2;
    }
    /* back substitution */
    if (norm != 0.0) {
        for (en = n; en >= 1; en--) {
            p = wr[en - 1];
            q = wi[en - 1];
            na = en - 1;
            if (q == 0.0) {
                m = en;
                H_(en - 1, en - 1) = 1.0;
                if (na != 0) {
                    for (i = en - 2; i >= 0; i--) {
                        w = H_(i, i) - p;
                        r = 0.0;
                        for (j = m - 1; j < en; j++) r += H_(i, j) * H_(j, en - 1);
                        if (wi[i] < 0.0) {
                            z = w;
                            s = r;
                        } else {
                            m = i + 1;
                            if (wi[i] == 0.0) {
                                t = w;
                                if (t == 0.0) {
                                    tst1 = norm;
                                    t = tst1;
                                    do {
                                        t = 0.01 * t;
                                        tst2 = norm + t;
                                    } while (tst2 > tst1);
                                }
                                H_(i, en - 1) = -(r / t);
                            } else {
                                x = H_(i, i + 1);
                                y = H_(i + 1, i);
                                q = (wr[i] - p) * (wr[i] - p) + wi[i] * wi[i];
                                t = (x * s - z * r) / q;
                                H_(i, en - 1) = t;
                                if (fabs(x) > fabs(z))
                                    H_(i + 1, en - 1) = (-r - w * t) / x;
                                else
                                    H_(i + 1, en - 1) = (-s - y * t) / z;
                            }
                            t = fabs(H_(i, en - 1));
                            if (t != 0.0) {
                                tst1 = t;
                                tst2 = tst1 + 1.0 / tst1;
                                if (tst2 <= tst1)
                                    for (j = i; j < en; j++) H_(j, en - 1) /= t;
                            }
                        }
                    }
                }
            } else if (q > 0.0) {
                m = na;
                if (fabs(H_(en - 1, na - 1)) > fabs(H_(na - 1, en - 1))) {
                    H_(na - 1, na - 1) = q / H_(en - 1, na - 1);
                    H_(na - 1, en - 1) = (p - H_(en - 1, en - 1)) / H_(en - 1, na - 1);
                } else {
                    mcdiv(0.0, -H_(na - 1, en - 1), H_(na - 1, na - 1) - p, q, c);
                    H_(na - 1, na - 1) = c[0];
                    H_(na - 1, en - 1) = c[1];
                }
                H_(en - 1, na - 1) = 0.0;
                H_(en - 1, en - 1) = 1.0;
                if (en != 2) {
                    for (i = en - 3; i >= 0; i--) {
                        w = H_(i, i) - p;
                        ra = 0.0;
                        sa = 0.0;
                        for (j = m - 1; j < en; j++) {
                            ra += H_(i, j) * H_(j, na - 1);
                            sa += H_(i, j) * H_(j, en - 1);
                        }
                        if (wi[i] < 0.0) {
                            z = w;
                            r = ra;
                            s = sa;
                        } else {
                            m = i + 1;
                            if (wi[i] == 0.0) {
                                mcdiv(-ra, -sa, w, q, c);
                                H_(i, na - 1) = c[0];
                                H_(i, en - 1) = c[1];
                            } else {
                                x = H_(i, i + 1);
                                y = H_(i + 1, i);
                                vr = (wr[i] - p) * (wr[i] - p);
                                vr = vr + wi[i] * wi[i] - q * q;
                                vi = (wr[i] - p) * 2.0 * q;
                                if (vr == 0.0 && vi == 0.0) {
                                    tst1 = norm * (fabs(w) + fabs(q) + fabs(x) + fabs(y) + fabs(z));
                                    vr = tst1;
                                    do {
                                        vr = 0.01 * vr;
                                        tst2 = tst1 + vr;
                                    } while (tst2 > tst1);
                                }
                                mcdiv(x * r - z * ra + q * sa, x * s - z * sa - q * ra, vr, vi, c);
                                H_(i, na - 1) = c[0];
                                H_(i, en - 1) = c[1];
                                if (fabs(x) > fabs(z) + fabs(q)) {
                                    H_(i + 1, na - 1) = (q * H_(i, en - 1) - w * H_(i, na - 1) - ra) / x;
                                    H_(i + 1, en - 1) = (-sa - w * H_(i, en - 1) - q * H_(i, na - 1)) / x;
                                } else {
                                    mcdiv(-r - y * H_(i, na - 1), -s - y * H_(i, en - 1), z, q, c);
                                    H_(i + 1, na - 1) = c[0];
                                    H_(i + 1, en - 1) = c[1];
                                }
                            }
                            t = (fabs(H_(i, na - 1)) > fabs(H_(i, en - 1))) ? fabs(H_(i, na - 1))
                                                                           : fabs(H_(i, en - 1));
                            if (t != 0.0) {
                                tst1 = t;
                                tst2 = tst1 + 1.0 / tst1;
                                if (tst2 <= tst1) {
                                    for (j = i; j < en; j++) {
                                        H_(j, na - 1) /= t;
                                        H_(j, en - 1) /= t;
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        for (i = 0; i < n; i++) {
            if (i + 1 < low || i + 1 > hgh)
                for (j = i; j < n; j++) Z_(i, j) = H_(i, j);
        }
        for (j = n - 1; j >= low - 1; j--) {
            m = ((j + 1) < hgh) ? (j + 1) : hgh;
            for (i = low - 1; i < hgh; i++) {
                z = 0.0;
                for (k = low - 1; k < m; k++) z += Z_(i, k) * H_(k, j);
                Z_(i, j) = z;
            }
        }
    }
    return 0;
}

/* Matrix.java:733-837 */
static int luinverse(const double *inmat, double *imtrx, int size)
{
    const double EPSILON = 2.220446049250313E-16;
    int i, j, k, l, maxi = 0, idx, ix, jx, n = size;
    double sum, tmp, maxb, aw;
    int index[OM_MAXK];
    double wk[OM_MAXK];
    double om[OM_MAXK * OM_MAXK];
#define O_(i, j) om[(i) * n + (j)]
    for (i = 0; i < size; i++)
        for (j = 0; j < size; j++) O_(i, j) = inmat[i * n + j];
    aw = 1.0;
    for (i = 0; i < size; i++) {
        maxb = 0.0;
        for (j = 0; j < size; j++)
            if (fabs(O_(i, j)) > maxb) maxb = fabs(O_(i, j));
        if (maxb == 0.0) return -1;
        wk[i] = 1.0 / maxb;
    }
    for (j = 0; j < size; j++) {
        for (i = 0; i < j; i++) {
            sum = O_(i, j);
            for (k = 0; k < i; k++) sum -= O_(i, k) * O_(k, j);
            O_(i, j) = sum;
        }
        maxb = 0.0;
        for (i = j; i < size; i++) {
            sum = O_(i, j);
            for (k = 0; k < j; k++) sum -= O_(i, k) * O_(k, j);
            O_(i, j) = sum;
            tmp = wk[i] * fabs(sum);
            if (tmp >= maxb) {
                maxb = tmp;
                maxi = i;
            }
        }
        if (j != maxi) {
            for (k = 0; k < size; k++) {
                tmp = O_(maxi, k);
                O_(maxi, k) = O_(j, k);
                O_(j, k) = tmp;
            }
            aw = -aw;
            wk[maxi] = wk[j];
        }
        index[j] = maxi;
        if (O_(j, j) == 0.0) O_(j, j) = EPSILON;
        if (j != size - 1) {
            tmp = 1.0 / O_(j, j);
            for (i = j + 1; i < size; i++) O_(i, j) *= tmp;
        }
    }
    (void)aw;
    for (jx = 0; jx < size; jx++) {
        for (ix = 0; ix < size; ix++) wk[ix] = 0.0;
        wk[jx] = 1.0;
        l = -1;
        for (i = 0; i < size; i++) {
            idx = index[i];
            sum = wk[idx];
            wk[idx] = wk[i];
            if (l != -1) {
                for (j = l; j < i; j++) sum -= O_(i, j) * wk[j];
            } else if (sum != 0.0) {
                l = i;
            }
            wk[i] = sum;
        }
        for (i = size - 1; i >= 0; i--) {
            sum = wk[i];
            for (j = i + 1; j < size; j++) sum -= O_(i, j) * wk[j];
            wk[i] = sum / O_(i, i);
        }
        for (ix = 0; ix < size; ix++) imtrx[ix * n + jx] = wk[ix];
    }
#undef O_
    return 0;
}

/* Matrix.java:42-60 */
int om_eigen(int n, const double *A, double *evec, double *wr, double *wi, double *ievec)
{
    double res[OM_MAXK * OM_MAXK];
    int ordr[OM_MAXK];
    if (n <= 0 || n > OM_MAXK) return -2;
    memcpy(res, A, sizeof(double) * n * n);
    memset(evec, 0, sizeof(double) * n * n);
    elmhes(res, ordr, n);
    eltran(res, evec, ordr, n);
    if (hqr2(n, 1, n, res, evec, wr, wi) != 0) return -1;
    if (luinverse(evec, ievec, n) != 0) return -3;
    return 0;
}

/* ------------------------------------------------------------------ */
/* models                                                              */
/* ------------------------------------------------------------------ */
struct om_model {
    int K;
    double F[OM_MAXK], prior[OM_MAXK];
    double R[OM_MAXK * OM_MAXK];
    double evec[OM_MAXK * OM_MAXK], ievec[OM_MAXK * OM_MAXK];
    double eval[OM_MAXK], eval_i[OM_MAXK];
};

/* SubstModel.java:80-110, :192-219 */
om_model *om_model_create_custom(int K, const double *F, const double *M, int symmetric, int normalise)
{
    int i, j;
    om_model *m;
    if (K < 2 || K > OM_MAXK) return NULL;
    m = (om_model *)calloc(1, sizeof(om_model));
    if (!m) return NULL;
    m->K = K;
    for (i = 0; i < K; i++) m->F[i] = F[i];
#define R_(i, j) m->R[(i) * K + (j)]
    for (i = 0; i < K; i++) {
        if (symmetric) {
            for (j = i + 1; j < K; j++) {
                double s = M[i * K + j];
                R_(i, j) = s * F[j];
                R_(j, i) = s * F[i];
            }
        } else {
            for (j = 0; j < K; j++) {
                if (i == j) continue;
                R_(i, j) = M[i * K + j];
            }
        }
    }
    /* makeValid */
    for (i = 0; i < K; i++) {
        double sum = 0.0;
        for (j = 0; j < K; j++)
            if (i != j) sum += R_(i, j);
        R_(i, i) = -sum;
    }
    if (normalise) {
        double sum = 0.0;
        for (i = 0; i < K; i++) sum += -R_(i, i) * F[i];
        for (i = 0; i < K; i++)
            for (j = 0; j < K; j++) R_(i, j) = R_(i, j) / sum;
    }
#undef R_
    if (om_eigen(K, m->R, m->evec, m->eval, m->eval_i, m->ievec) != 0) {
        free(m);
        return NULL;
    }
    /* root prior: new EnumDistrib(domain, F) -> copy + normalise (EnumDistrib.java:66-78, :363-371) */
    {
        double sum = 0.0;
        for (i = 0; i < K; i++) sum += F[i];
        for (i = 0; i < K; i++) m->prior[i] = F[i] / sum;
    }
    return m;
}

static void unpack_upper(const double *S, double *M, int K)
{
    int i, j, k = 0;
    memset(M, 0, sizeof(double) * K * K);
    for (i = 0; i < K; i++)
        for (j = i + 1; j < K; j++) M[i * K + j] = S[k++];
}

om_model *om_model_create(const char *name, const double *F_override)
{
    double M[400];
    const double *F = NULL, *S = NULL;
    if (!strcmp(name, "LG")) { F = BNK_LG_F; S = BNK_LG_S; }
    else if (!strcmp(name, "JTT")) { F = BNK_JTT_F; S = BNK_JTT_S; }
    else if (!strcmp(name, "WAG")) { F = BNK_WAG_F; S = BNK_WAG_S; }
    else if (!strcmp(name, "Dayhoff")) { F = BNK_DAYHOFF_F; S = BNK_DAYHOFF_S; }
    if (F) {
        unpack_upper(S, M, 20);
        return om_model_create_custom(20, F_override ? F_override : F, M, 1, 1);
    }
    if (!strcmp(name, "Yang")) /* Yang.java:32-50: Q as published, normalised */
        return om_model_create_custom(4, F_override ? F_override : BNK_YANG_F, BNK_YANG_Q, 0, 1);
    if (!strcmp(name, "JC")) { /* JC.java:32-57: F = 1/n, Q off-diagonal = mu/n, mu = 1, NOT normalised */
        double Fj[4], Qj[16];
        int r, c;
        for (r = 0; r < 4; r++) {
            Fj[r] = 1.0 / 4;
            for (c = 0; c < 4; c++) Qj[r * 4 + c] = (r != c) ? 1.0 / 4 : 0.0;
        }
        return om_model_create_custom(4, Fj, Qj, 0, 0);
    }
    return NULL;
}

void om_model_free(om_model *m) { free(m); }
int om_model_nstates(const om_model *m) { return m->K; }

void om_model_get(const om_model *m, double *R, double *evec, double *ievec, double *eval,
                  double *F, double *prior)
{
    int K = m->K;
    if (R) memcpy(R, m->R, sizeof(double) * K * K);
    if (evec) memcpy(evec, m->evec, sizeof(double) * K * K);
    if (ievec) memcpy(ievec, m->ievec, sizeof(double) * K * K);
    if (eval) memcpy(eval, m->eval, sizeof(double) * K);
    if (F) memcpy(F, m->F, sizeof(double) * K);
    if (prior) memcpy(prior, m->prior, sizeof(double) * K);
}

/* SubstModel.java:303-327 */
void om_getprobs(const om_model *m, double time, double *P)
{
    int i, j, k, K = m->K;
    double temp;
    double iexp[OM_MAXK * OM_MAXK];
    for (k = 0; k < K; k++) {
        temp = om_exp(time * m->eval[k]);
        for (j = 0; j < K; j++) iexp[k * K + j] = m->ievec[k * K + j] * temp;
    }
    for (i = 0; i < K; i++) {
        for (j = 0; j < K; j++) {
            temp = 0.0;
            for (k = 0; k < K; k++) temp += m->evec[i * K + k] * iexp[k * K + j];
            P[i * K + j] = fabs(temp);
        }
    }
}

/* ------------------------------------------------------------------ */
/* forest construction (IdxTree.java:411-464)                          */
/* ------------------------------------------------------------------ */
int om_prune(int n, const int *parent, const uint8_t *present_in, int remove_orphans,
             uint8_t *present_out, int *n_roots)
{
    int i, cnt = 0, roots = 0;
    /* pruneMe = { i : !present_in[i] } */
    if (remove_orphans) {
        uint8_t *orphan = (uint8_t *)malloc(n);
        uint8_t *isleaf = (uint8_t *)malloc(n);
        int *stack = (int *)malloc(sizeof(int) * n);
        /* child lists in index order (IdxTree.children[]) */
        int *nch = (int *)calloc(n + 1, sizeof(int));
        int *first = (int *)malloc(sizeof(int) * (n + 1));
        int *kids = (int *)malloc(sizeof(int) * n);
        memset(orphan, 1, n);
        memset(isleaf, 1, n);
        for (i = 0; i < n; i++)
            if (parent[i] >= 0) { isleaf[parent[i]] = 0; nch[parent[i]]++; }
        first[0] = 0;
        for (i = 0; i < n; i++) first[i + 1] = first[i] + nch[i];
        memset(nch, 0, sizeof(int) * (n + 1));
        for (i = 0; i < n; i++)
            if (parent[i] >= 0) kids[first[parent[i]] + nch[parent[i]]++] = i;
        for (i = 0; i < n; i++) {
            if (!isleaf[i]) continue;
            int top = -1, idx = i;
            while (present_in[idx] && orphan[idx]) {
                orphan[idx] = 0;
                top = idx;
                idx = parent[idx];
                if (idx == -1) break;
            }
            if (top != -1) {
                /* markMyChildrenAsNonOrphans(top): flood present descendants (iterative) */
                int sp = 0;
                stack[sp++] = top;
                while (sp > 0) {
                    int v = stack[--sp], c;
                    if (isleaf[v] || !present_in[v]) continue;
                    for (c = first[v]; c < first[v + 1]; c++) {
                        int u = kids[c];
                        if (!orphan[u] || !present_in[u]) continue;
                        orphan[u] = 0;
                        stack[sp++] = u;
                    }
                }
            }
        }
        for (i = 0; i < n; i++) present_out[i] = (uint8_t)(present_in[i] && !orphan[i]);
        free(orphan); free(isleaf); free(stack); free(nch); free(first); free(kids);
    } else {
        for (i = 0; i < n; i++) present_out[i] = present_in[i] ? 1 : 0;
    }
    for (i = 0; i < n; i++) {
        if (!present_out[i]) continue;
        cnt++;
        if (parent[i] < 0 || !present_out[parent[i]]) roots++;
    }
    if (n_roots) *n_roots = roots;
    return cnt;
}
