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

// pairs a group of ncols columns can hold: distinct (state, state) codes of two children, a multiple of 4
__host__ __device__ inline int cherry_cap(int K, int ncols)
{
    const int c = (K + 1) * (K + 1) < ncols ? (K + 1) * (K + 1) : ncols;
    return (c + 3) & ~3;
}

constexpr int CG = 2048;  // columns per cherry group (api.cu builds groups of one category, <= CG columns)

// Result table of one (node y, group x) in a global scratch area (L2-resident between the two kernels), state-major
// so that the expansion's gathers of one state row fall into few cache lines: K x cap message values, K x cap
// traceback pointers, cap rescale exponents, where cap = the group's capacity in pairs (a multiple of 4).  The
// table starts at byte ((y * recs_per_node + group.ncat) * REC_BYTES) -- api.cu stores the group's record offset
// in TileDesc::ncat and sizes it with cherry_group_capacity().
template <int K>
struct CherryTable {
    static constexpr int REC_BYTES = ((K * 9 + 4 + 15) / 16) * 16;
    double *v;
    uint8_t *ptr;
    int32_t *e;
    int cap;
    __device__ __forceinline__ CherryTable(uint8_t *scratch, size_t node_row, int recs_per_node, const TileDesc &g)
    {
        cap = cherry_cap(K, g.ncols);
        uint8_t *base = scratch + ((size_t)node_row * recs_per_node + g.ncat) * REC_BYTES;
        v = reinterpret_cast<double *>(base);
        ptr = base + (size_t)K * cap * 8;
        e = reinterpret_cast<int32_t *>(base + (size_t)K * cap * 9);
    }
};

// "Cherry-direct" (r2): a height-1 node under a level-scheduled parent never has its message rows written out at all.
// Its consumers -- the parent's up-sweep thread, the parent's fused down-sweep thread, the traceback -- read the
// column's record index (a.cslots) and gather the record from the (node, group) table the pairs kernel left in the
// scratch area: 20 loads of 8 bytes out of a few KB that stay in L1 / L2 while a CTA walks the tiles of one group,
// instead of 160 bytes per (node, column) written to HBM by an expansion kernel and read back twice.
// The group of a column: a.tile_group[tile] names the first group overlapping the tile; a tile straddles at most
// one group boundary (groups of a split category hold >= 1,024 columns).
__device__ __forceinline__ TileDesc cherry_group_of(const WideArgs &a, int tile_index, int col)
{
    int gi = a.tile_group[tile_index];
    TileDesc g = a.cgroups[gi];
    if (col >= g.col0 + g.ncols) g = a.cgroups[gi + 1];
    return g;
}
// message of the height-1 node in row y of its level for column col, all K states (MODE as in the level kernels)
template <int K>
__device__ __forceinline__ void cherry_gather(const WideArgs &a, int y, int tile_index, int col, double (&g)[K])
{
    const TileDesc grp = cherry_group_of(a, tile_index, col);
    const CherryTable<K> tab(a.cscratch, (size_t)y, a.crecs_per_node, grp);
    const int r = a.cslots[(size_t)y * a.ncols + col];
    const double *v = tab.v + r;
#pragma unroll
    for (int x = 0; x < K; x++) g[x] = __ldg(v + (size_t)x * tab.cap);
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
