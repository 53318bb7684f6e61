// sweeps_fast.cu -- lean variants of the three tree sweeps for the common GRASP case:
// every node present in every column (no indel mask), evidence only on childless nodes
// (alignment leaves), at most two children per node, message stack entirely in shared
// memory.  Same arithmetic and association order as sweeps.cu (which stays the general
// path for masks, internal evidence, multifurcations and stack spill); what differs is the
// instruction budget: the per-node control flow is branch-free, addresses are per-thread
// base pointers plus one multiply-add, and nothing per-column is tested twice.
//
// Profile that motivated it (profiles/r1_ncu_v2_summary.md): the general kernel retired
// ~400 instructions per warp per node, ~125 of them the 20x20 contraction itself; it was
// issue-bound (63 % issue-slot utilisation, FP64 pipe 13 %, DRAM < 3 % of peak).
#include "device_common.cuh"

#include <cstddef>

namespace bnk {

constexpr int FSTAGE = 8;   // table ring depth (power of two)
constexpr int FRING = 64;   // walk-program ring: two 32-step chunks ...
constexpr int FCHUNK = 32;
constexpr int FGUARD = 8;   // ... plus a mirror of its first steps, so look-ahead never wraps
// Depth of the per-thread register pipeline that feeds the childless children: the observed state of step
// i + F_OA is requested while step i runs, the table entry it selects at step i + F_LA.  A step of a deep tree
// takes ~300-500 cycles once nothing stalls, a DRAM round trip ~1,500: with the r1 depths (2 and 1) every step
// of the 10,000-level caterpillar waited for both loads (1.43 us per step, profiles/r1_v12_bench_caterpillar).
constexpr int F_LA = 4;
constexpr int F_OA = 8;
static_assert(F_OA <= FGUARD && F_LA < F_OA, "look-ahead must stay inside the program ring's guard");

__device__ __forceinline__ void f_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void f_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

size_t fast_smem_bytes(int K, int cpt, int ncg, int slots)
{
    const size_t lrow = K + 2, tcpad = (size_t)ncg * cpt;
    // program ring | table ring (one category per tile) | gs x2 | unit row | stack (+1 dummy slot)
    return (FRING + FGUARD) * sizeof(Step) +
           sizeof(double) * ((size_t)FSTAGE * K * lrow + 2 * tcpad * lrow + lrow + (size_t)(slots + 1) * tcpad * K);
}

// Shared-memory layout (byte offsets from the dynamic shared base).  Everything is addressed
// as smem_raw + offset so that the compiler keeps 32-bit shared addressing and hoists the base.
template <int K>
struct FLayout {
    static constexpr int LROW = K + 2;
    static constexpr int PROG = 0;
    static constexpr int TAB = (FRING + FGUARD) * (int)sizeof(Step);
    static constexpr int STAGE_BYTES = K * LROW * 8;
    static constexpr int GS = TAB + FSTAGE * STAGE_BYTES;
    int gs_bytes, unit, stack, slot_bytes;
    __device__ FLayout(int tcpad)
    {
        gs_bytes = tcpad * LROW * 8;
        unit = GS + 2 * gs_bytes;
        stack = unit + LROW * 8;
        slot_bytes = tcpad * K * 8;
    }
};

__device__ __forceinline__ void f_cp16(unsigned smem_addr, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gsrc) : "memory");
}

// per-thread plan for streaming one K x K table (K*K/2 16-byte pieces) into a padded stage
template <int K>
struct TableCopy {
    static constexpr int NCH = K * K / 2;
    int src_off, dst_off;  // bytes, for piece q = threadIdx.x (valid when threadIdx.x < NCH)
    __device__ TableCopy()
    {
        constexpr int LROW = K + 2, CH = K / 2;
        const int q = threadIdx.x < NCH ? threadIdx.x : 0;
        const int row = q / CH, ch = q - row * CH;
        src_off = q * 16;
        dst_off = (row * LROW + ch * 2) * 8;
    }
    __device__ __forceinline__ void issue(const char *table, unsigned stage_addr) const
    {
        constexpr int LROW = K + 2, CH = K / 2;
        if (threadIdx.x < NCH) f_cp16(stage_addr + dst_off, table + src_off);
        if ((int)blockDim.x < NCH) {  // small CTAs: the remaining pieces
            for (int q = threadIdx.x + blockDim.x; q < NCH; q += blockDim.x) {
                const int row = q / CH, ch = q - row * CH;
                f_cp16(stage_addr + (row * LROW + ch * 2) * 8, table + q * 16);
            }
        }
    }
};

__device__ __forceinline__ void f_issue_prog(const Step *prog_g, int n_steps, int chunk, unsigned ring_addr)
{
    const unsigned dst = ring_addr + (chunk & 1) * FCHUNK * (unsigned)sizeof(Step);
    for (int k = threadIdx.x; k < FCHUNK * 4; k += blockDim.x) {
        const int s = chunk * FCHUNK + (k >> 2);
        if (s < n_steps) {
            const char *src = reinterpret_cast<const char *>(prog_g + s) + (k & 3) * 16;
            f_cp16(dst + k * 16, src);
            if ((chunk & 1) == 0 && k < FGUARD * 4) f_cp16(ring_addr + FRING * (unsigned)sizeof(Step) + k * 16, src);
        }
    }
}

__device__ __forceinline__ double f_pow2_scale(int mhi, int &e)
{
    // mhi = largest high word among non-negative doubles; returns 2^-e, e = unbiased exponent of that maximum
    const int be = mhi >> 20;
    if (be == 0) { e = 0; return 1.0; }
    e = be - 1023;
    return __hiloint2double((2046 - be) << 20, 0);
}

// 1/s to ~1 ulp without the IEEE division slow path (posteriors are checked to 1e-9)
__device__ __forceinline__ double f_rcp(double s)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
    r = fma(r, fma(-s, r, 1.0), r);
    r = fma(r, fma(-s, r, 1.0), r);
    return r;
}

#define STEP_FIELD(i, member) \
    (*reinterpret_cast<const int32_t *>(smem_raw + (((i) & (FRING - 1)) * (int)sizeof(Step)) + (int)offsetof(Step, member)))
// same, from the running ring offset po of step i and a look-ahead of k steps (k <= FGUARD)
#define STEP_AT(po, k, member) \
    (*reinterpret_cast<const int32_t *>(smem_raw + (po) + (k) * (int)sizeof(Step) + (int)offsetof(Step, member)))
#define SM_D(off) (*reinterpret_cast<double *>(smem_raw + (off)))
#define SM_D2(off) (*reinterpret_cast<const double2 *>(smem_raw + (off)))

// ------------------------------------------------------------------------------------
// up-sweeps.  MODE 0: max-product, log space (joint).  MODE 1: sum-product, linear space.
//
// Threads beyond the tile's columns shadow a valid column of the tile and perform the same
// loads, arithmetic and (identical-value) stores: the loop body has no "is this thread
// active" branch at all.
// ------------------------------------------------------------------------------------
template <int K, int CPT, int MODE>
__global__ void __launch_bounds__(CPT == 1 ? 512 : (CPT == 2 ? 384 : 256)) up_fast_kernel(const WalkArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using LY = FLayout<K>;
    constexpr int LROW = K + 2, KK = K * K;
    const TileDesc tile = a.tiles[blockIdx.x];
    const int tcpad = a.ncg * CPT;
    const LY ly(tcpad);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem_raw);
    const int tid = threadIdx.x, p = tid % K, cg = tid / K;
    const int n_steps = a.n_steps, ncols = a.ncols;
    const unsigned ncols_u = (unsigned)ncols;
    const double ident = (MODE == 0) ? 0.0 : 1.0;
    const int dummy_off = ly.stack + a.smem_slots * ly.slot_bytes;  // one spare slot behind the real ones
    const char *tab_cat = reinterpret_cast<const char *>(a.T_row + (size_t)tile.cat0 * a.n_nodes * KK);
    const double *tcol_cat_p = a.T_col + (size_t)tile.cat0 * a.n_nodes * KK + p;
    const TableCopy<K> tcopy;

    int sidx[CPT], gsw[CPT], ocol[CPT], scol[CPT];
    const uint8_t *obs_c[CPT];
    uint8_t *ptr_o[CPT];
    double *msg_o[CPT];
    const double *wmsg_p[CPT];
#pragma unroll
    for (int j = 0; j < CPT; j++) {
        int lc = cg * CPT + j;
        if (cg >= a.ncg) lc = j;                                      // spare threads shadow group 0
        const int lcc = lc < tile.ncols ? lc : tile.ncols - 1;        // spare columns shadow the last column
        const int col = tile.col0 + lcc;
        ocol[j] = a.orig_col[col];
        sidx[j] = (lc * K + p) * 8;
        gsw[j] = lc * LROW * 8;
        obs_c[j] = a.obs + ocol[j];  // inputs stay in the caller's column order
        scol[j] = col;
        ptr_o[j] = (MODE == 0) ? a.ptr + (size_t)col * K + p : nullptr;
        msg_o[j] = (MODE == 1 && a.msg) ? a.msg + (size_t)col * K + p : nullptr;
        wmsg_p[j] = a.msg ? a.msg + (size_t)p * ncols + col : nullptr;  // wide layout [ent][K][ncols]
    }
    const unsigned row_stride = (unsigned)(ncols * K);  // ptr / msg rows (host guarantees ncols * K < 2^31)
    const double prior_p = a.prior[p];
    const int rowoff = p * LROW * 8;
    for (int q = tid; q < LROW; q += blockDim.x) SM_D(ly.unit + q * 8) = ident;

    int esum[CPT];
    double logacc[CPT];
#pragma unroll
    for (int j = 0; j < CPT; j++) { esum[j] = 0; logacc[j] = 0.0; }

    f_issue_prog(a.up, n_steps, 0, sbase + LY::PROG);
    f_cp_async_commit();
    f_cp_async_wait<0>();
    __syncthreads();
    for (int s = 0; s < FSTAGE - 1; s++) {
        if (s < n_steps) tcopy.issue(tab_cat + (size_t)STEP_FIELD(s, node) * (KK * 8), sbase + LY::TAB + s * LY::STAGE_BYTES);
        f_cp_async_commit();
    }

    // pipeline registers (rings shifted once per step): obr[k] = observed state of the child at step
    // i + F_LA + k, gr[k] = its value at step i + k, isobs bit (2k + e) = "child e of step i + k is observed"
    constexpr int NOB = F_OA - F_LA + 1, NG = F_LA + 1;
    int obr[CPT][2][NOB];
    double gr[CPT][2][NG];
    unsigned isobs[CPT];
    auto load_obs = [&](int po, int k, int j, int e) -> int {
        const int cn = (e == 0) ? STEP_AT(po, k, c_node[0]) : STEP_AT(po, k, c_node[1]);
        return obs_c[j][(size_t)(unsigned)(cn < 0 ? 0 : cn) * ncols_u];
    };
    auto load_look = [&](int po, int k, int ob, int j, int e) -> double {
        const int cn = (e == 0) ? STEP_AT(po, k, c_node[0]) : STEP_AT(po, k, c_node[1]);
        const int cs = (e == 0) ? STEP_AT(po, k, c_slot[0]) : STEP_AT(po, k, c_slot[1]);
        const int ce = (e == 0) ? STEP_AT(po, k, c_ent[0]) : STEP_AT(po, k, c_ent[1]);
        const bool childless = cn >= 0 && cs == SLOT_CHILDLESS, wide = cs == SLOT_WIDE;
        const int node = cn < 0 ? 0 : cn;
        // one load either way: the table entry of an observed childless child, or the message a
        // wide level kernel left in HBM ([ent][K][ncols])
        const double *src = wide ? wmsg_p[j] + (size_t)(unsigned)ce * row_stride
                                 : tcol_cat_p + (unsigned)(node * KK + (ob == BNK_NONE ? 0 : ob) * K);
        double g = *src;
        if (!childless && !wide) g = ident;
        else if (childless && ob == BNK_NONE) {  // uninstantiated childless node: maximise / marginalise its own row
            const double *row = a.T_row + ((size_t)tile.cat0 * a.n_nodes + node) * KK + p * K;
            g = row[0];
            for (int x = 1; x < K; x++) g = (MODE == 0) ? (row[x] > g ? row[x] : g) : g + row[x];
        }
        return g;
    };
#pragma unroll
    for (int j = 0; j < CPT; j++) {
        isobs[j] = 0;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            int o[F_OA];
#pragma unroll
            for (int k = 0; k < F_OA; k++) o[k] = k < n_steps ? load_obs(0, k, j, e) : BNK_NONE;
#pragma unroll
            for (int k = 0; k < F_LA; k++) {
                gr[j][e][k] = k < n_steps ? load_look(0, k, o[k], j, e) : ident;
                isobs[j] |= (unsigned)(o[k] != BNK_NONE) << (2 * k + e);
            }
#pragma unroll
            for (int k = 0; k < NOB - 1; k++) obr[j][e][k] = o[F_LA + k];
        }
    }

    for (int i = 0; i < n_steps; i++) {
        const int po = (i & (FRING - 1)) * (int)sizeof(Step);
        const int ent = STEP_AT(po, 0, ent), slot = STEP_AT(po, 0, slot);
        const int cs0 = STEP_AT(po, 0, c_slot[0]), cs1 = STEP_AT(po, 0, c_slot[1]);
        const bool rootstep = STEP_AT(po, 0, parent) < 0;
        const int gs_off = LY::GS + (i & 1) * ly.gs_bytes;
        // ---- pipeline: states for step i + F_OA, table entries for step i + F_LA ----
        {
            const bool ho = i + F_OA < n_steps, hl = i + F_LA < n_steps;
#pragma unroll
            for (int j = 0; j < CPT; j++) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    obr[j][e][NOB - 1] = ho ? load_obs(po, F_OA, j, e) : BNK_NONE;
                    gr[j][e][NG - 1] = hl ? load_look(po, F_LA, obr[j][e][0], j, e) : ident;
                    isobs[j] |= (unsigned)(obr[j][e][0] != BNK_NONE) << (2 * F_LA + e);
                }
            }
        }
        // ---- fold the (at most two) children; identities make the formula uniform ----
        const bool in0 = cs0 >= 0, in1 = cs1 >= 0;
        const int so0 = in0 ? ly.stack + cs0 * ly.slot_bytes : dummy_off;
        const int so1 = in1 ? ly.stack + cs1 * ly.slot_bytes : dummy_off;
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            const double s0 = SM_D(so0 + sidx[j]), s1 = SM_D(so1 + sidx[j]);
            double c0 = in0 ? s0 : gr[j][0][0], c1 = in1 ? s1 : gr[j][1][0];
            if (MODE == 0 && rootstep) {
                // a root combines [prior, observed children, unobserved children] (the order the factors
                // sit in the reference's bucket); with one child of each kind that fixes the association
                const bool o0 = cs0 == SLOT_CHILDLESS && (isobs[j] & 1u);
                const bool o1 = cs1 == SLOT_CHILDLESS && (isobs[j] & 2u);
                if (!o0 && o1) { const double t = c0; c0 = c1; c1 = t; }
            }
            double G;
            if (MODE == 0) G = ((rootstep ? prior_p : 0.0) + c0) + c1;
            else G = ((rootstep ? prior_p : 1.0) * c0) * c1;
            SM_D(gs_off + gsw[j] + p * 8) = G;
        }
        f_cp_async_wait<FSTAGE - 2>();
        __syncthreads();
        if (i + FSTAGE - 1 < n_steps)
            tcopy.issue(tab_cat + (size_t)(unsigned)STEP_AT(po, FSTAGE - 1, node) * (KK * 8),
                        sbase + LY::TAB + ((i + FSTAGE - 1) & (FSTAGE - 1)) * LY::STAGE_BYTES);
        if ((i & (FCHUNK - 1)) == 0) f_issue_prog(a.up, n_steps, (i >> 5) + 1, sbase + LY::PROG);
        f_cp_async_commit();
        // ---- contract with the node's table ----
        const int row_off = rootstep ? ly.unit : LY::TAB + (i & (FSTAGE - 1)) * LY::STAGE_BYTES + rowoff;
        const int so = slot >= 0 ? ly.stack + slot * ly.slot_bytes : dummy_off;
        if (MODE == 0) {
            double best[CPT];
            int bi[CPT];
#pragma unroll
            for (int j = 0; j < CPT; j++) { best[j] = -CUDART_INF; bi[j] = 0; }
#pragma unroll
            for (int x2 = 0; x2 < K / 2; x2++) {
                const double2 l = SM_D2(row_off + x2 * 16);
#pragma unroll
                for (int j = 0; j < CPT; j++) {
                    const double2 g = SM_D2(gs_off + gsw[j] + x2 * 16);
                    const double v0 = l.x + g.x, v1 = l.y + g.y;
                    const bool t0 = v0 > best[j];
                    best[j] = t0 ? v0 : best[j];
                    bi[j] = t0 ? 2 * x2 : bi[j];
                    const bool t1 = v1 > best[j];
                    best[j] = t1 ? v1 : best[j];
                    bi[j] = t1 ? 2 * x2 + 1 : bi[j];
                }
            }
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                SM_D(so + sidx[j]) = best[j];
                ptr_o[j][(size_t)(unsigned)ent * row_stride] = (uint8_t)bi[j];
            }
        } else {
            double acc0[CPT], acc1[CPT];
            int mhi[CPT];
#pragma unroll
            for (int j = 0; j < CPT; j++) { acc0[j] = 0.0; acc1[j] = 0.0; mhi[j] = 0; }
#pragma unroll
            for (int x2 = 0; x2 < K / 2; x2++) {
                const double2 l = SM_D2(row_off + x2 * 16);
#pragma unroll
                for (int j = 0; j < CPT; j++) {
                    const double2 g = SM_D2(gs_off + gsw[j] + x2 * 16);
                    acc0[j] = fma(l.x, g.x, acc0[j]);
                    acc1[j] = fma(l.y, g.y, acc1[j]);
                    mhi[j] = max(mhi[j], max(__double2hiint(g.x), __double2hiint(g.y)));
                }
            }
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                int e;
                const double scale = f_pow2_scale(mhi[j], e);
                const double m = (acc0[j] + acc1[j]) * scale;
                esum[j] += e;
                SM_D(so + sidx[j]) = m;
                if (rootstep) logacc[j] += log(m);
                else if (msg_o[j]) msg_o[j][(size_t)(unsigned)ent * row_stride] = m;
            }
        }
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            isobs[j] >>= 2;
#pragma unroll
            for (int e = 0; e < 2; e++) {
#pragma unroll
                for (int k = 0; k < NOB - 1; k++) obr[j][e][k] = obr[j][e][k + 1];
#pragma unroll
                for (int k = 0; k < NG - 1; k++) gr[j][e][k] = gr[j][e][k + 1];
            }
        }
    }
    if (MODE == 1 && a.out_logl && p == 0) {
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            const int ew = a.esum_wide ? a.esum_wide[scol[j]] : 0;
            a.out_logl[ocol[j]] = logacc[j] + (double)(esum[j] + ew) * 0.693147180559945309417232121458;
        }
    }
}

// ------------------------------------------------------------------------------------
// down-sweep (see sweeps.cu for the recurrences)
// ------------------------------------------------------------------------------------
template <int K, int CPT>
__global__ void __launch_bounds__(CPT == 1 ? 512 : (CPT == 2 ? 384 : 256)) down_fast_kernel(const WalkArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using LY = FLayout<K>;
    constexpr int LROW = K + 2, KK = K * K;
    const TileDesc tile = a.tiles[blockIdx.x];
    const int tcpad = a.ncg * CPT;
    const LY ly(tcpad);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem_raw);
    const int tid = threadIdx.x, p = tid % K, cg = tid / K;
    const int n_steps = a.n_steps, ncols = a.ncols;
    const unsigned ncols_u = (unsigned)ncols;
    const int dummy_off = ly.stack + a.smem_slots * ly.slot_bytes;
    const char *tab_cat = reinterpret_cast<const char *>(a.T_row + (size_t)tile.cat0 * a.n_nodes * KK);
    const double *tcol_cat_p = a.T_col + (size_t)tile.cat0 * a.n_nodes * KK + p;
    const TableCopy<K> tcopy;

    int sidx[CPT], gsw[CPT], srd[CPT];
    const uint8_t *obs_c[CPT];
    const double *msg_i[CPT];
    const double *wmsg_p[CPT];
    double *tmsg_p[CPT];
    double *post_o[CPT];
#pragma unroll
    for (int j = 0; j < CPT; j++) {
        int lc = cg * CPT + j;
        if (cg >= a.ncg) lc = j;
        const int lcc = lc < tile.ncols ? lc : tile.ncols - 1;
        const int col = tile.col0 + lcc;
        sidx[j] = (lc * K + p) * 8;
        srd[j] = lc * K * 8;
        gsw[j] = lc * LROW * 8;
        obs_c[j] = a.obs + a.orig_col[col];  // inputs stay in the caller's column order
        msg_i[j] = a.msg + (size_t)col * K + p;
        wmsg_p[j] = a.msg + (size_t)p * ncols + col;                      // wide layout [ent][K][ncols]
        tmsg_p[j] = a.tmsg ? a.tmsg + (size_t)p * ncols + col : nullptr;  // likewise
        post_o[j] = a.out_post ? a.out_post + (size_t)a.orig_col[col] * K + p : nullptr;
    }
    const unsigned row_stride = (unsigned)(ncols * K);
    const double prior_p = a.prior[p];
    const int rowoff = p * LROW * 8;

    f_issue_prog(a.down, n_steps, 0, sbase + LY::PROG);
    f_cp_async_commit();
    f_cp_async_wait<0>();
    __syncthreads();
    // ring: step i uses stage i & 3; the table of step i + FSTAGE is requested right after the
    // barrier of step i
    for (int s = 0; s < FSTAGE; s++) {
        if (s < n_steps) tcopy.issue(tab_cat + (size_t)STEP_FIELD(s, node) * (KK * 8), sbase + LY::TAB + s * LY::STAGE_BYTES);
        f_cp_async_commit();
    }
    constexpr int NOB = F_OA - F_LA + 1, NG = F_LA + 1;
    int obr[CPT][2][NOB];  // rings as in up_fast_kernel
    double gr[CPT][2][NG];
    int qr[NG];            // posterior row of steps i .. i + F_LA
    auto load_obs = [&](int po, int k, int j, int e) -> int {
        const int cn = (e == 0) ? STEP_AT(po, k, c_node[0]) : STEP_AT(po, k, c_node[1]);
        return obs_c[j][(size_t)(unsigned)(cn < 0 ? 0 : cn) * ncols_u];
    };
    // value of child e: stored up-message (internal child), table entry (observed childless child),
    // row sum (uninstantiated childless child), 1 (no such child)
    auto load_val = [&](int po, int k, int ob, int j, int e) -> double {
        const int cn = (e == 0) ? STEP_AT(po, k, c_node[0]) : STEP_AT(po, k, c_node[1]);
        const int ce = (e == 0) ? STEP_AT(po, k, c_ent[0]) : STEP_AT(po, k, c_ent[1]);
        const int cs = (e == 0) ? STEP_AT(po, k, c_slot[0]) : STEP_AT(po, k, c_slot[1]);
        const bool exists = cn >= 0, internal = ce >= 0, wide = cs == SLOT_WIDE;
        const int node = cn < 0 ? 0 : cn;
        const double *src = wide ? wmsg_p[j] + (size_t)(unsigned)ce * row_stride
                          : internal ? msg_i[j] + (size_t)(unsigned)ce * row_stride
                                     : tcol_cat_p + (unsigned)(node * KK + (ob == BNK_NONE ? 0 : ob) * K);
        double g = *src;
        if (!exists) g = 1.0;
        else if (!internal && ob == BNK_NONE) {
            g = 0.0;
            for (int x = 0; x < K; x++) g += tcol_cat_p[(size_t)node * KK + x * K];
        }
        return g;
    };
#pragma unroll
    for (int j = 0; j < CPT; j++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            int o[F_OA];
#pragma unroll
            for (int k = 0; k < F_OA; k++) o[k] = k < n_steps ? load_obs(0, k, j, e) : BNK_NONE;
#pragma unroll
            for (int k = 0; k < F_LA; k++) gr[j][e][k] = k < n_steps ? load_val(0, k, o[k], j, e) : 1.0;
#pragma unroll
            for (int k = 0; k < NOB - 1; k++) obr[j][e][k] = o[F_LA + k];
        }
    }
#pragma unroll
    for (int k = 0; k < F_LA; k++) qr[k] = k < n_steps ? (a.qrow ? a.qrow[STEP_FIELD(k, ent)] : STEP_FIELD(k, ent)) : -1;
    f_cp_async_wait<FSTAGE - 1>();
    __syncthreads();

    for (int i = 0; i < n_steps; i++) {
        const int po = (i & (FRING - 1)) * (int)sizeof(Step);
        const int slot = STEP_AT(po, 0, slot);
        const bool rootstep = STEP_AT(po, 0, parent) < 0;
        const int dcs0 = STEP_AT(po, 0, c_slot[0]), dcs1 = STEP_AT(po, 0, c_slot[1]);
        const bool in0 = dcs0 >= 0, in1 = dcs1 >= 0;
        const bool wd0 = dcs0 == SLOT_WIDE, wd1 = dcs1 == SLOT_WIDE;
        const unsigned we0 = (unsigned)STEP_AT(po, 0, c_ent[0]), we1 = (unsigned)STEP_AT(po, 0, c_ent[1]);
        const int so0 = in0 ? ly.stack + dcs0 * ly.slot_bytes : dummy_off;
        const int so1 = in1 ? ly.stack + dcs1 * ly.slot_bytes : dummy_off;
        const int gs_off = LY::GS + (i & 1) * ly.gs_bytes;
        {
            const bool ho = i + F_OA < n_steps, hl = i + F_LA < n_steps;
#pragma unroll
            for (int j = 0; j < CPT; j++) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    obr[j][e][NOB - 1] = ho ? load_obs(po, F_OA, j, e) : BNK_NONE;
                    gr[j][e][NG - 1] = hl ? load_val(po, F_LA, obr[j][e][0], j, e) : 1.0;
                }
            }
            qr[NG - 1] = hl ? (a.qrow ? a.qrow[STEP_AT(po, F_LA, ent)] : STEP_AT(po, F_LA, ent)) : -1;
        }
        // ---- from-above vector: D = P^T T (scaled by a power of two), or the prior at a root ----
        const int row_off = LY::TAB + (i & (FSTAGE - 1)) * LY::STAGE_BYTES + rowoff;
        const int ts = slot >= 0 ? ly.stack + slot * ly.slot_bytes : dummy_off;
        double acc0[CPT], acc1[CPT], U[CPT];
        int mhi[CPT];
#pragma unroll
        for (int j = 0; j < CPT; j++) { acc0[j] = 0.0; acc1[j] = 0.0; mhi[j] = 0; }
#pragma unroll
        for (int x2 = 0; x2 < K / 2; x2++) {
            const double2 l = SM_D2(row_off + x2 * 16);
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                const double2 t = SM_D2(ts + srd[j] + x2 * 16);
                acc0[j] = fma(l.x, t.x, acc0[j]);
                acc1[j] = fma(l.y, t.y, acc1[j]);
                mhi[j] = max(mhi[j], max(__double2hiint(t.x), __double2hiint(t.y)));
            }
        }
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            int e;
            const double scale = f_pow2_scale(mhi[j], e);
            const double D = rootstep ? prior_p : (acc0[j] + acc1[j]) * scale;
            const double c0 = gr[j][0][0], c1 = gr[j][1][0];
            U[j] = D * (c0 * c1);
            SM_D(gs_off + gsw[j] + p * 8) = U[j];
            // child slots never alias this step's own slot (tree_plan.cpp)
            SM_D(so0 + sidx[j]) = D * c1;
            SM_D(so1 + sidx[j]) = D * c0;
            // children handled by the wide level kernels receive their from-above vector through HBM
            if (wd0) tmsg_p[j][(size_t)we0 * row_stride] = D * c1;
            if (wd1) tmsg_p[j][(size_t)we1 * row_stride] = D * c0;
        }
        f_cp_async_wait<FSTAGE - 2>();
        __syncthreads();
        if (i + FSTAGE < n_steps)
            tcopy.issue(tab_cat + (size_t)(unsigned)STEP_AT(po, FSTAGE, node) * (KK * 8),
                        sbase + LY::TAB + (i & (FSTAGE - 1)) * LY::STAGE_BYTES);
        if ((i & (FCHUNK - 1)) == 0) f_issue_prog(a.down, n_steps, (i >> 5) + 1, sbase + LY::PROG);
        f_cp_async_commit();
        // ---- normalise and emit the posterior ----
        const int q0 = qr[0];
        if (q0 >= 0 && a.out_post) {
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                double s = 0.0;
#pragma unroll
                for (int x2 = 0; x2 < K / 2; x2++) { const double2 g = SM_D2(gs_off + gsw[j] + x2 * 16); s += g.x; s += g.y; }
                post_o[j][(size_t)(unsigned)q0 * row_stride] = U[j] * f_rcp(s);
            }
        }
#pragma unroll
        for (int k = 0; k < NG - 1; k++) qr[k] = qr[k + 1];
#pragma unroll
        for (int j = 0; j < CPT; j++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
#pragma unroll
                for (int k = 0; k < NOB - 1; k++) obr[j][e][k] = obr[j][e][k + 1];
#pragma unroll
                for (int k = 0; k < NG - 1; k++) gr[j][e][k] = gr[j][e][k + 1];
            }
        }
    }
}

// ------------------------------------------------------------------------------------
template <typename KernelT>
static cudaError_t launch_fast(KernelT kern, const WalkLaunch &wl, const WalkArgs &a)
{
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wl.smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<wl.n_tiles, wl.threads, wl.smem_bytes, wl.stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_joint_up_fast(const WalkLaunch &wl, const WalkArgs &a)
{
    if (wl.K == 20) {
        if (wl.cpt == 1) return launch_fast(up_fast_kernel<20, 1, 0>, wl, a);
        if (wl.cpt == 2) return launch_fast(up_fast_kernel<20, 2, 0>, wl, a);
        if (wl.cpt == 4) return launch_fast(up_fast_kernel<20, 4, 0>, wl, a);
    } else if (wl.K == 4 && wl.cpt == 1) return launch_fast(up_fast_kernel<4, 1, 0>, wl, a);
    return cudaErrorInvalidValue;
}
cudaError_t launch_sum_up_fast(const WalkLaunch &wl, const WalkArgs &a)
{
    if (wl.K == 20) {
        if (wl.cpt == 1) return launch_fast(up_fast_kernel<20, 1, 1>, wl, a);
        if (wl.cpt == 2) return launch_fast(up_fast_kernel<20, 2, 1>, wl, a);
        if (wl.cpt == 4) return launch_fast(up_fast_kernel<20, 4, 1>, wl, a);
    } else if (wl.K == 4 && wl.cpt == 1) return launch_fast(up_fast_kernel<4, 1, 1>, wl, a);
    return cudaErrorInvalidValue;
}
cudaError_t launch_sum_down_fast(const WalkLaunch &wl, const WalkArgs &a)
{
    if (wl.K == 20) {
        if (wl.cpt == 1) return launch_fast(down_fast_kernel<20, 1>, wl, a);
        if (wl.cpt == 2) return launch_fast(down_fast_kernel<20, 2>, wl, a);
        if (wl.cpt == 4) return launch_fast(down_fast_kernel<20, 4>, wl, a);
    } else if (wl.K == 4 && wl.cpt == 1) return launch_fast(down_fast_kernel<4, 1>, wl, a);
    return cudaErrorInvalidValue;
}

}  // namespace bnk
