// wide_common.cuh -- helpers shared by the level kernels (wide.cu: lean path, wide_general.cu: masks / evidence).
#pragma once
#include "device_common.cuh"

namespace bnk {

constexpr int WT = 128;  // columns (= threads) per CTA

template <int K>
__device__ __forceinline__ void load_table(double *dst, const double *src, int pad)
{
    // dst rows of K + pad doubles
    for (int q = threadIdx.x; q < K * K; q += WT) dst[(q / K) * (K + pad) + (q % K)] = src[q];
}

__device__ __forceinline__ double w_pow2_scale(int mhi, int &e)
{
    const int be = mhi >> 20;
    if (be == 0) { e = 0; return 1.0; }
    e = be - 1023;
    return __hiloint2double((2046 - be) << 20, 0);
}

__device__ __forceinline__ double w_rcp(double s)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
    r = fma(r, fma(-s, r, 1.0), r);
    r = fma(r, fma(-s, r, 1.0), r);
    return r;
}

// One candidate of the max-product scan: v = l + g; if (v > best) { best = v; bi = x; }  (strict '>',
// so the first maximum in ascending x wins, Factorize.java:1472-1481).  The update re-issues the add
// under the predicate instead of moving a 64-bit value with two selects: 4 issue slots per candidate
// (DADD, DSETP, @P DADD, @P MOV) instead of 5 -- the kernel is issue-bound, the FP64 pipe has headroom.
__device__ __forceinline__ void argmax_step(double &best, int &bi, double l, double g, int x)
{
    asm("{\n\t.reg .pred p;\n\t.reg .f64 v;\n\t"
        "add.rn.f64 v, %2, %3;\n\t"
        "setp.gt.f64 p, v, %0;\n\t"
        "@p add.rn.f64 %0, v, 0d0000000000000000;\n\t"
        "@p mov.s32 %1, %4;\n\t}"
        : "+d"(best), "+r"(bi)
        : "d"(l), "d"(g), "r"(x));
}

}  // namespace bnk
