// wide_common.cuh -- helpers shared by the level kernels (wide.cu: lean path, wide_general.cu: masks / evidence).
#pragma once
#include "device_common.cuh"

namespace bnk {

constexpr int WT = 128;  // columns (= threads) per CTA

// Stage a K x K table into shared memory rows of K + pad doubles with asynchronous 8-byte copies (LDGSTS): every
// copy of every table is in flight before anything waits, so a CTA pays one memory round trip for its tables
// instead of one per loop iteration (ncu on the first cherry kernel: long-scoreboard stalls 14.5 per issue).
// Call tables_wait() before the __syncthreads() that publishes them.
template <int K>
__device__ __forceinline__ void load_table(double *dst, const double *src, int pad)
{
#pragma unroll
    for (int q0 = 0; q0 < K * K; q0 += WT) {
        const int q = q0 + threadIdx.x;
        if (q < K * K) {
            const unsigned s = (unsigned)__cvta_generic_to_shared(dst + (q / K) * (K + pad) + (q % K));
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(src + q) : "memory");
        }
    }
}
__device__ __forceinline__ void tables_wait() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

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

// One posterior row (K doubles of one (query, column), contiguous) leaves as K / 4 256-bit streaming stores
// (STG.256, sm_100): a thread then writes whole 32-byte sectors.  With 128-bit stores every sector was written
// by two instructions, i.e. twice the L2 write requests for the same bytes -- and the level kernels sit closer
// to the L2's request rate than to anything else (tools/microbench/layout_bw.cu: a kernel that only moves these
// rows reaches the measured 6.5 TB/s).  Rows that are not 32-byte aligned (a caller's odd pointer) fall back.
template <int K>
__device__ __forceinline__ void store_post_row(double *row, const double (&U)[K], double r, bool aligned32)
{
    if (K % 4 == 0 && aligned32) {
#pragma unroll
        for (int x4 = 0; x4 < K / 4; x4++)
            asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};\n" ::"l"(row + 4 * x4), "d"(U[4 * x4] * r), "d"(U[4 * x4 + 1] * r),
                         "d"(U[4 * x4 + 2] * r), "d"(U[4 * x4 + 3] * r)
                         : "memory");
    } else {
        double2 *o = reinterpret_cast<double2 *>(row);
#pragma unroll
        for (int x2 = 0; x2 < K / 2; x2++) __stcs(o + x2, make_double2(U[2 * x2] * r, U[2 * x2 + 1] * r));
    }
}

// Max-out of one table row (Factorize.getMaxMargin, src/bn/factor/Factorize.java:1472-1481): the first strict
// maximum of v[x] = Lrow[x] + G[x] scanning x ascending from (-inf, index 0).  Run as a tournament in which a
// tie keeps the lower index: the winner is the lowest index holding the maximal value -- the same answer as the
// sequential scan (and index 0 with -inf when every candidate is -inf), with a 5-deep dependency chain instead
// of 20 and 4.7 issue slots per candidate (DADD, DSETP, 2 FSEL, SEL) instead of the 6 ptxas makes of a
// predicated re-add (measured on B200, tools/microbench/joint_scan.cu: +16 % updates/s).
template <int K>
__device__ __forceinline__ void argmax_row(const double *Lrow, const double (&G)[K], double &best, int &bi)
{
    static_assert(K % 2 == 0, "rows are read as double2");
    constexpr int C = 8;  // candidates per sub-tournament: bounds the live values (register pressure)
    const double2 *row2 = reinterpret_cast<const double2 *>(Lrow);
#pragma unroll
    for (int c0 = 0; c0 < K; c0 += C) {
        constexpr int dummy = 0;
        (void)dummy;
        const int n = (K - c0 < C) ? K - c0 : C;
        double v[C];
        int idx[C];
#pragma unroll
        for (int x2 = 0; x2 < C / 2; x2++) {
            if (2 * x2 < n) {
                const double2 l = row2[c0 / 2 + x2];
                v[2 * x2] = l.x + G[c0 + 2 * x2];
                v[2 * x2 + 1] = l.y + G[c0 + 2 * x2 + 1];
                idx[2 * x2] = c0 + 2 * x2;
                idx[2 * x2 + 1] = c0 + 2 * x2 + 1;
            }
        }
#pragma unroll
        for (int s = 1; s < C; s *= 2) {
#pragma unroll
            for (int x = 0; x + s < C; x += 2 * s) {
                if (x + s < n) {
                    const bool gt = v[x + s] > v[x];
                    v[x] = gt ? v[x + s] : v[x];
                    idx[x] = gt ? idx[x + s] : idx[x];
                }
            }
        }
        if (c0 == 0) {
            best = v[0];
            bi = idx[0];
        } else {
            const bool gt = v[0] > best;
            best = gt ? v[0] : best;
            bi = gt ? idx[0] : bi;
        }
    }
}

// An uninstantiated childless node maximises (MODE 0) / marginalises (MODE 1) its own table rows: g[x] over
// the K rows of the table at L (global memory; rare case).
template <int K, int MODE>
__device__ __forceinline__ void self_vector(const double *L, double (&g)[K])
{
#pragma unroll
    for (int x = 0; x < K; x++) {
        double r = L[x * K];
        for (int y = 1; y < K; y++) r = (MODE == 0) ? (L[x * K + y] > r ? L[x * K + y] : r) : r + L[x * K + y];
        g[x] = r;
    }
}

}  // namespace bnk
