// wide.cu -- level kernels: all nodes of one height level in one launch, one COLUMN per thread.
//
// Where a height level of the tree holds many nodes (tree_plan.cpp: >= 16), every (node, column)
// pair of the level is independent, so the level runs as one launch over a (column tiles x nodes)
// grid -- the "level-order schedule" of BASELINE.json's north star.  A thread owns one column:
// the 20-state vector it combines lives in registers, the node's 20x20 table sits in shared
// memory and is read as warp-uniform broadcasts (one wavefront serves 32 columns), leaf children
// are table lookups by the observed state, and messages travel through HBM in
// [node][state][column] layout so that every load and store is a coalesced row.
// This is the HBM-bound formulation of SURVEY.md 8(d); the column-owned walker (sweeps_fast.cu)
// keeps the deep, narrow top of the tree, reading these messages as its leaf-level input.
//
// Same arithmetic as the walker (and the oracle): joint adds in child order then L_v[p][x] + G[x],
// first strict maximum scanning x ascending; sum-product in linear space with an exact 2^-e
// rescale per (node, column) whose exponents are summed into the log-likelihood.
// Preconditions (checked by the host): nothing masked, evidence on childless nodes only, at most
// two children per node, one rate category per column tile, wide nodes are never roots.
#include "wide_common.cuh"

namespace bnk {

// message of child e of node w for this thread's column, all K states.  MODE 0: log space, 1: linear.
template <int K, int MODE>
__device__ __forceinline__ void child_vector(const WideArgs &a, const WideNode &w, int e, size_t tb, int col, bool valid,
                                             const double *lt_smem, double (&g)[K], int tile_index = 0)
{
    const int cn = w.c_node[e];
    const int crow = (w.c_ent[e] >= 0 && a.cherry_row) ? a.cherry_row[w.c_ent[e]] : -1;
    if (crow >= 0) {  // a height-1 child whose message was never expanded: gather its record (cherry-direct)
        cherry_gather<K>(a, crow, tile_index, col, g);
    } else if (w.c_ent[e] >= 0) {  // internal child: its message row-block in HBM
        const double *m = a.msg + (size_t)w.c_ent[e] * K * a.ncols + col;
#pragma unroll
        for (int x = 0; x < K; x++) g[x] = m[(size_t)x * a.ncols];
    } else {
        const int ou = a.obs[(size_t)cn * a.ncols + a.orig_col[col]];  // inputs stay in the caller's column order
        if (ou != BNK_NONE) {
#pragma unroll
            for (int x = 0; x < K; x++) g[x] = lt_smem[ou * (K + 1) + x];
        } else {  // uninstantiated childless node: maximise / marginalise its own row
            self_vector<K, MODE>(a.T_row + (tb + cn) * (K * K), g);
        }
    }
    (void)valid;
}

// ------------------------------------------------------------------------------------
// up-sweeps
// ------------------------------------------------------------------------------------
template <int K, int MODE>
__global__ void __launch_bounds__(WT, 5) up_wide_kernel(const WideArgs a)
{
    __shared__ __align__(16) double Ls[K * K];
    __shared__ double LT0[K * (K + 1)], LT1[K * (K + 1)];
    const WideNode w = a.nodes[blockIdx.y];
    const int t_lo = blockIdx.x * a.tiles_per_cta, t_hi = min(t_lo + a.tiles_per_cta, a.n_tiles);
    int cur_cat = -1;
    for (int ti = t_lo; ti < t_hi; ti++) {
        const TileDesc tile = a.tiles[ti];
        const size_t tb = (size_t)tile.cat0 * a.n_nodes;
        if (tile.cat0 != cur_cat) {  // uniform: tables of this node for the tile's rate category
            __syncthreads();
            load_table<K>(Ls, a.T_row + (tb + w.node) * (K * K), 0);
            if (w.n_child > 0 && w.c_ent[0] < 0) load_table<K>(LT0, a.T_col + (tb + w.c_node[0]) * (K * K), 1);
            if (w.n_child > 1 && w.c_ent[1] < 0) load_table<K>(LT1, a.T_col + (tb + w.c_node[1]) * (K * K), 1);
            tables_wait();
            __syncthreads();
            cur_cat = tile.cat0;
        }
        const bool valid = (int)threadIdx.x < tile.ncols;
        const int col = tile.col0 + (valid ? (int)threadIdx.x : tile.ncols - 1);

        double G[K], g[K];
        child_vector<K, MODE>(a, w, 0, tb, col, valid, LT0, G, ti);
        if (w.n_child > 1) {
            child_vector<K, MODE>(a, w, 1, tb, col, valid, LT1, g, ti);
#pragma unroll
            for (int x = 0; x < K; x++) G[x] = (MODE == 0) ? G[x] + g[x] : G[x] * g[x];
        }
        double *mout = a.msg + (size_t)w.ent * K * a.ncols + col;
        if (MODE == 0) {
            uint8_t *pout = a.ptr + (size_t)w.ent * K * a.ncols + col;
#pragma unroll 2
            for (int p = 0; p < K; p++) {
                double best;
                int bi;
                argmax_row<K>(Ls + p * K, G, best, bi);
                if (valid) {
                    mout[(size_t)p * a.ncols] = best;
                    pout[(size_t)p * a.ncols] = (uint8_t)bi;
                }
            }
        } else {
            int mhi = 0;
#pragma unroll
            for (int x = 0; x < K; x++) mhi = max(mhi, __double2hiint(G[x]));
            int e;
            const double scale = w_pow2_scale(mhi, e);
#pragma unroll
            for (int x = 0; x < K; x++) G[x] *= scale;  // exact: power of two
#pragma unroll 2
            for (int p = 0; p < K; p++) {
                const double2 *row2 = reinterpret_cast<const double2 *>(Ls + p * K);
                double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
                for (int x2 = 0; x2 < K / 2; x2++) {
                    const double2 l = row2[x2];
                    acc0 = fma(l.x, G[2 * x2], acc0);
                    acc1 = fma(l.y, G[2 * x2 + 1], acc1);
                }
                if (valid) mout[(size_t)p * a.ncols] = acc0 + acc1;
            }
            if (valid && a.escale) a.escale[(size_t)(a.wide_row0 + blockIdx.y) * a.ncols + col] = (int16_t)e;
        }
    }
}

// ------------------------------------------------------------------------------------
// down-sweep: T (from above, HBM) -> D = P^T T -> posterior, and the children's from-above vectors
// ------------------------------------------------------------------------------------
// CHERRY: every node of the level has only childless children (height 1): nothing is pushed down,
// far fewer live registers, more resident warps.
template <int K, bool CHERRY>
__global__ void __launch_bounds__(WT, CHERRY ? 4 : 3) down_wide_kernel(const WideArgs a)
{
    __shared__ __align__(16) double PTs[K * K];
    __shared__ double LT0[K * (K + 1)], LT1[K * (K + 1)];
    const WideNode w = a.nodes[blockIdx.y];
    const int t_lo = blockIdx.x * a.tiles_per_cta, t_hi = min(t_lo + a.tiles_per_cta, a.n_tiles);
    const bool two = w.n_child > 1;
    const bool leaf0 = CHERRY || w.c_ent[0] < 0, leaf1 = two && (CHERRY || w.c_ent[1] < 0);
    const bool push0 = !leaf0, push1 = two && !leaf1;
    const int q = a.qrow ? a.qrow[w.ent] : w.ent;
    const bool post32 = (reinterpret_cast<uintptr_t>(a.out_post) & 31) == 0;
    int cur_cat = -1;
    for (int ti = t_lo; ti < t_hi; ti++) {
        const TileDesc tile = a.tiles[ti];
        const size_t tb = (size_t)tile.cat0 * a.n_nodes;
        if (tile.cat0 != cur_cat) {
            __syncthreads();
            load_table<K>(PTs, a.T_col + (tb + w.node) * (K * K), 0);  // PT[x][xp] = P[xp][x]
            if (leaf0) load_table<K>(LT0, a.T_col + (tb + w.c_node[0]) * (K * K), 1);
            if (leaf1) load_table<K>(LT1, a.T_col + (tb + w.c_node[1]) * (K * K), 1);
            tables_wait();
            __syncthreads();
            cur_cat = tile.cat0;
        }
        const bool valid = (int)threadIdx.x < tile.ncols;
        const int col = tile.col0 + (valid ? (int)threadIdx.x : tile.ncols - 1);

        // children: an observed leaf is a row of its P^T table (shared memory), an internal child
        // its up-message in HBM; uninstantiated leaves (rare) go through child_vector
        int o0 = BNK_NONE, o1 = BNK_NONE;
        const int ocol = a.orig_col[col];  // inputs stay in the caller's column order
        if (leaf0) o0 = a.obs[(size_t)w.c_node[0] * a.ncols + ocol];
        if (leaf1) o1 = a.obs[(size_t)w.c_node[1] * a.ncols + ocol];
        const bool slow = (leaf0 && o0 == BNK_NONE) || (leaf1 && o1 == BNK_NONE);
        const double *m0 = leaf0 ? nullptr : a.msg + (size_t)w.c_ent[0] * K * a.ncols + col;
        const double *m1 = (two && !leaf1) ? a.msg + (size_t)w.c_ent[1] * K * a.ncols + col : nullptr;
        double *t0 = push0 ? a.tmsg + (size_t)w.c_ent[0] * K * a.ncols + col : nullptr;
        double *t1 = push1 ? a.tmsg + (size_t)w.c_ent[1] * K * a.ncols + col : nullptr;
        double T[K];
        {
            const double *t = a.tmsg + (size_t)w.ent * K * a.ncols + col;
            int mhi = 0;
#pragma unroll
            for (int x = 0; x < K; x++) { T[x] = t[(size_t)x * a.ncols]; }
#pragma unroll
            for (int x = 0; x < K; x++) mhi = max(mhi, __double2hiint(T[x]));
            int e;
            const double scale = w_pow2_scale(mhi, e);
#pragma unroll
            for (int x = 0; x < K; x++) T[x] *= scale;
        }
        double U[K], S = 0.0;
        if (!slow) {
            // phase 1: D = P^T T for all states (registers only)
            double D[K];
#pragma unroll
            for (int x = 0; x < K; x++) {
                const double2 *row2 = reinterpret_cast<const double2 *>(PTs + x * K);
                double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
                for (int y2 = 0; y2 < K / 2; y2++) {
                    const double2 l = row2[y2];
                    acc0 = fma(l.x, T[2 * y2], acc0);
                    acc1 = fma(l.y, T[2 * y2 + 1], acc1);
                }
                D[x] = acc0 + acc1;
            }
            // phase 2: every child value is requested before anything is stored (the stores below may
            // alias these loads as far as the compiler knows, which would serialise them)
            double c0[K], c1[K];
#pragma unroll
            for (int x = 0; x < K; x++) {
                c0[x] = leaf0 ? LT0[o0 * (K + 1) + x] : __ldg(m0 + (size_t)x * a.ncols);
                c1[x] = !two ? 1.0 : (leaf1 ? LT1[o1 * (K + 1) + x] : __ldg(m1 + (size_t)x * a.ncols));
            }
            // phase 3: from-above vectors of internal children, unnormalised posterior
#pragma unroll
            for (int x = 0; x < K; x++) {
                if (valid && push0) t0[(size_t)x * a.ncols] = D[x] * c1[x];
                if (valid && push1) t1[(size_t)x * a.ncols] = D[x] * c0[x];
                U[x] = D[x] * (c0[x] * c1[x]);
                S += U[x];
            }
        } else {
            double g0[K], g1[K];
            child_vector<K, 1>(a, w, 0, tb, col, valid, LT0, g0);
            if (two) child_vector<K, 1>(a, w, 1, tb, col, valid, LT1, g1);
            else {
#pragma unroll
                for (int x = 0; x < K; x++) g1[x] = 1.0;
            }
#pragma unroll
            for (int x = 0; x < K; x++) {
                double D = 0.0;
                for (int y = 0; y < K; y++) D = fma(PTs[x * K + y], T[y], D);
                if (valid && push0) t0[(size_t)x * a.ncols] = D * g1[x];
                if (valid && push1) t1[(size_t)x * a.ncols] = D * g0[x];
                U[x] = D * (g0[x] * g1[x]);
                S += U[x];
            }
        }
        if (valid && q >= 0 && a.out_post) {
            // each thread writes its column's 160 contiguous bytes (staging the tile through shared
            // memory for wider stores was measured: no gain, two extra barriers)
            store_post_row<K>(a.out_post + ((size_t)q * a.ncols + ocol) * K, U, w_rcp(S), post32);
        }
    }
}

// ------------------------------------------------------------------------------------
// down-sweep of a level FUSED with its height-1 children (r2).  A height-1 node ("cherry") has only
// childless children, so once its parent v knows its own from-above product D_v, the cherry's posterior is one
// more 20 x 20 contraction away: T_c = D_v * g_sibling, D_c = P_c^T T_c, post_c ~ D_c * (leaf rows).  Done in the
// parent's thread, T_c never travels through HBM (160 B written + 160 B read per cherry and column, 20 % of the
// down-sweep's DRAM traffic on a balanced tree) and the height-1 level needs no launch of its own.  Taller
// children get their from-above vector through HBM as in down_wide_kernel.  Same arithmetic
// and order as down_wide_kernel on both levels (incl. the power-of-two rescale of T_c); results agree to a few ulp
// (the compiler contracts the multiply-adds of the two kernels differently).
// Uninstantiated leaves need no slow path here: row K of a staged leaf table holds the sum of its rows (what
// self_vector<K, 1> computes, same order), and BNK_NONE selects that row.
// ------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(WT, 3) down_wide2_kernel(const WideArgs a)  // (two CTAs per SM at 252 registers, no spills: 6.95 against 6.41 ms)
{
    constexpr int LS = K + 1, LTAB = LS * LS;
    __shared__ __align__(16) double PTv[K * K];
    __shared__ __align__(16) double PTc[2][K * K];
    __shared__ double LTv[2][LTAB], LTc[2][2][LTAB];
    const Wide2Node w2 = a.fuse2[blockIdx.y];
    const WideNode w = w2.v;
    const int t_lo = blockIdx.x * a.tiles_per_cta, t_hi = min(t_lo + a.tiles_per_cta, a.n_tiles);
    const bool two = w.n_child > 1;
    const bool ch0 = w2.c[0].node >= 0, ch1 = two && w2.c[1].node >= 0;   // children fused into this thread
    const bool leaf0 = w.c_ent[0] < 0, leaf1 = two && w.c_ent[1] < 0;
    const bool push0 = !leaf0 && !ch0, push1 = two && !leaf1 && !ch1;     // taller children: from-above vector through HBM
    const int q = a.qrow ? a.qrow[w.ent] : w.ent;
    const int qc0 = ch0 ? (a.qrow ? a.qrow[w2.c[0].ent] : w2.c[0].ent) : -1;
    const int qc1 = ch1 ? (a.qrow ? a.qrow[w2.c[1].ent] : w2.c[1].ent) : -1;
    const bool post32 = (reinterpret_cast<uintptr_t>(a.out_post) & 31) == 0;
    int cur_cat = -1;
    for (int ti = t_lo; ti < t_hi; ti++) {
        const TileDesc tile = a.tiles[ti];
        const size_t tb = (size_t)tile.cat0 * a.n_nodes;
        if (tile.cat0 != cur_cat) {
            __syncthreads();
            load_table<K>(PTv, a.T_col + (tb + w.node) * (K * K), 0);
            double *sumtab[4] = {nullptr, nullptr, nullptr, nullptr};
            int nt = 0;
            if (leaf0) { load_table<K>(LTv[0], a.T_col + (tb + w.c_node[0]) * (K * K), 1); sumtab[nt++] = LTv[0]; }
            if (leaf1) { load_table<K>(LTv[1], a.T_col + (tb + w.c_node[1]) * (K * K), 1); sumtab[nt++] = LTv[1]; }
#pragma unroll
            for (int e = 0; e < 2; e++) {
                if (e == 0 ? ch0 : ch1) {
                    const WideNode &c = w2.c[e];
                    load_table<K>(PTc[e], a.T_col + (tb + c.node) * (K * K), 0);
                    load_table<K>(LTc[e][0], a.T_col + (tb + c.c_node[0]) * (K * K), 1); sumtab[nt++] = LTc[e][0];
                    if (c.n_child > 1) { load_table<K>(LTc[e][1], a.T_col + (tb + c.c_node[1]) * (K * K), 1); sumtab[nt++] = LTc[e][1]; }
                }
            }
            tables_wait();
            __syncthreads();
            {   // row K of every leaf table = sum of its rows (an uninstantiated leaf marginalises its own state)
                const int wid = threadIdx.x >> 5, x = threadIdx.x & 31;
                if (wid < nt && x < K) {
                    double *tab = sumtab[wid];
                    double r = tab[x];
                    for (int o = 1; o < K; o++) r += tab[o * LS + x];
                    tab[K * LS + x] = r;
                }
            }
            __syncthreads();
            cur_cat = tile.cat0;
        }
        const bool valid = (int)threadIdx.x < tile.ncols;
        const int col = tile.col0 + (valid ? (int)threadIdx.x : tile.ncols - 1);
        const int ocol = a.orig_col[col];

        double T[K];
        {
            const double *t = a.tmsg + (size_t)w.ent * K * a.ncols + col;
            int mhi = 0;
#pragma unroll
            for (int x = 0; x < K; x++) { T[x] = t[(size_t)x * a.ncols]; }
#pragma unroll
            for (int x = 0; x < K; x++) mhi = max(mhi, __double2hiint(T[x]));
            int e;
            const double scale = w_pow2_scale(mhi, e);
#pragma unroll
            for (int x = 0; x < K; x++) T[x] *= scale;
        }
        // observed states: of v's own childless children and of the fused cherries' children (BNK_NONE -> sum row)
        int o0 = 0, o1 = 0, oc[2][2] = {{0, 0}, {0, 0}};
        if (leaf0) { o0 = a.obs[(size_t)w.c_node[0] * a.ncols + ocol]; o0 = o0 == BNK_NONE ? K : o0; }
        if (leaf1) { o1 = a.obs[(size_t)w.c_node[1] * a.ncols + ocol]; o1 = o1 == BNK_NONE ? K : o1; }
#pragma unroll
        for (int e = 0; e < 2; e++)
            if (e == 0 ? ch0 : ch1) {
#pragma unroll
                for (int k = 0; k < 2; k++)
                    if (k < w2.c[e].n_child) {
                        const int o = a.obs[(size_t)w2.c[e].c_node[k] * a.ncols + ocol];
                        oc[e][k] = o == BNK_NONE ? K : o;
                    }
            }
        const double *m0 = leaf0 ? nullptr : a.msg + (size_t)w.c_ent[0] * K * a.ncols + col;
        const double *m1 = (two && !leaf1) ? a.msg + (size_t)w.c_ent[1] * K * a.ncols + col : nullptr;
        // D = P^T T (registers only), then the children's vectors
        double D[K];
#pragma unroll
        for (int x = 0; x < K; x++) {
            const double2 *row2 = reinterpret_cast<const double2 *>(PTv + x * K);
            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
            for (int y2 = 0; y2 < K / 2; y2++) {
                const double2 l = row2[y2];
                acc0 = fma(l.x, T[2 * y2], acc0);
                acc1 = fma(l.y, T[2 * y2 + 1], acc1);
            }
            D[x] = acc0 + acc1;
        }
        double c0[K], c1[K];
        if (ch0 && a.cherry_row) cherry_gather<K>(a, w2.crow[0], ti, col, c0);  // cherry-direct: the record, not an HBM row
        else {
#pragma unroll
            for (int x = 0; x < K; x++) c0[x] = leaf0 ? LTv[0][o0 * LS + x] : __ldg(m0 + (size_t)x * a.ncols);
        }
        if (ch1 && a.cherry_row) cherry_gather<K>(a, w2.crow[1], ti, col, c1);
        else {
#pragma unroll
            for (int x = 0; x < K; x++) c1[x] = !two ? 1.0 : (leaf1 ? LTv[1][o1 * LS + x] : __ldg(m1 + (size_t)x * a.ncols));
        }
        // posterior of v
        {
            double S = 0.0;
#pragma unroll
            for (int x = 0; x < K; x++) S += D[x] * (c0[x] * c1[x]);
            if (valid && q >= 0 && a.out_post) {
                const double r = w_rcp(S);
                double *row = a.out_post + ((size_t)q * a.ncols + ocol) * K;
                if (post32) {
#pragma unroll
                    for (int x4 = 0; x4 < K / 4; x4++)
                        asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};\n" ::"l"(row + 4 * x4),
                                     "d"((D[4 * x4] * (c0[4 * x4] * c1[4 * x4])) * r), "d"((D[4 * x4 + 1] * (c0[4 * x4 + 1] * c1[4 * x4 + 1])) * r),
                                     "d"((D[4 * x4 + 2] * (c0[4 * x4 + 2] * c1[4 * x4 + 2])) * r), "d"((D[4 * x4 + 3] * (c0[4 * x4 + 3] * c1[4 * x4 + 3])) * r)
                                     : "memory");
                } else {
                    double2 *o = reinterpret_cast<double2 *>(row);
#pragma unroll
                    for (int x2 = 0; x2 < K / 2; x2++)
                        __stcs(o + x2, make_double2((D[2 * x2] * (c0[2 * x2] * c1[2 * x2])) * r,
                                                    (D[2 * x2 + 1] * (c0[2 * x2 + 1] * c1[2 * x2 + 1])) * r));
                }
            }
        }
        // from-above vectors of the two children, in place: c1 <- T_child0 = D * c1, c0 <- T_child1 = D * c0
#pragma unroll
        for (int x = 0; x < K; x++) {
            const double d = D[x], a0 = c0[x], a1 = c1[x];
            c1[x] = d * a1;
            c0[x] = d * a0;
        }
        if (valid && push0) {
            double *t0 = a.tmsg + (size_t)w.c_ent[0] * K * a.ncols + col;
#pragma unroll
            for (int x = 0; x < K; x++) t0[(size_t)x * a.ncols] = c1[x];
        }
        if (valid && push1) {
            double *t1 = a.tmsg + (size_t)w.c_ent[1] * K * a.ncols + col;
#pragma unroll
            for (int x = 0; x < K; x++) t1[(size_t)x * a.ncols] = c0[x];
        }
        // the fused cherries
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const bool fused = e == 0 ? ch0 : ch1;
            const int qc = e == 0 ? qc0 : qc1;
            if (!fused || qc < 0 || !a.out_post) continue;  // uniform
            double (&Tc)[K] = e == 0 ? c1 : c0;
            {
                int mhi = 0;
#pragma unroll
                for (int x = 0; x < K; x++) mhi = max(mhi, __double2hiint(Tc[x]));
                int ee;
                const double scale = w_pow2_scale(mhi, ee);
#pragma unroll
                for (int x = 0; x < K; x++) Tc[x] *= scale;
            }
            const bool ctwo = w2.c[e].n_child > 1;
            const double *ra = LTc[e][0] + oc[e][0] * LS, *rb = LTc[e][1] + oc[e][1] * LS;
            double U[K], S = 0.0;
#pragma unroll
            for (int x = 0; x < K; x++) {
                const double2 *row2 = reinterpret_cast<const double2 *>(PTc[e] + x * K);
                double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
                for (int y2 = 0; y2 < K / 2; y2++) {
                    const double2 l = row2[y2];
                    acc0 = fma(l.x, Tc[2 * y2], acc0);
                    acc1 = fma(l.y, Tc[2 * y2 + 1], acc1);
                }
                U[x] = (acc0 + acc1) * (ra[x] * (ctwo ? rb[x] : 1.0));
                S += U[x];
            }
            if (valid) store_post_row<K>(a.out_post + ((size_t)qc * a.ncols + ocol) * K, U, w_rcp(S), post32);
        }
    }
}

// ------------------------------------------------------------------------------------
// traceback of one level (processed from the highest wide level down): one (node, column) per thread
// ------------------------------------------------------------------------------------
// (a thread keeps its column and walks a slice of the level's nodes, four dependent byte chains in flight)
__global__ void __launch_bounds__(256) traceback_wide_kernel(const WideArgs a, int K, int n_nodes)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col < a.col_lo || col >= a.col_hi) return;
#pragma unroll 4
    for (int r = blockIdx.y; r < n_nodes; r += gridDim.y) {
        const int32_t parent = a.nodes[r].parent, node = a.nodes[r].node, ent = a.nodes[r].ent;
        const int sp = a.state_s[(size_t)parent * a.ncols + col];
        a.state_s[(size_t)node * a.ncols + col] = a.ptr[((size_t)ent * K + sp) * a.ncols + col];
    }
}

// sum of the per-node rescale exponents of the wide part: esum[col] += sum over a slab of rows
__global__ void sum_escale_kernel(const int16_t *__restrict__ escale, int n_rows, int ncols, int32_t *__restrict__ esum)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    const int r0 = blockIdx.y * 128, r1 = min(r0 + 128, n_rows);
    int s = 0;
    for (int r = r0; r < r1; r++) s += escale[(size_t)r * ncols + col];
    atomicAdd(esum + col, s);
}

// out[row(node)][orig_col[c]] = in[row(node)][c] for the listed nodes (sorted -> original column order)
// (a thread keeps its column and walks a slice of the rows: one CTA per (row, 256 columns) was 200,000 CTAs of one
// byte per thread at C3 -- 0.165 ms for 50 MB)
__global__ void scatter_rows_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                                    const int32_t *__restrict__ nodes, int n_rows, const int32_t *__restrict__ orig_col,
                                    int ncols)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    const int oc = orig_col[c];
#pragma unroll 4
    for (int r = blockIdx.y; r < n_rows; r += gridDim.y) {
        const size_t row = (size_t)nodes[r] * ncols;
        out[row + oc] = in[row + c];
    }
}

// ------------------------------------------------------------------------------------
cudaError_t launch_up_wide(int K, int mode, const WideArgs &a, int n_tiles, int n_nodes, cudaStream_t st)
{
    if (n_nodes == 0 || n_tiles == 0) return cudaSuccess;
    if (a.cherry && a.cgroups) return launch_up_cherry(K, mode, a, n_nodes, st);  // site-pattern compression (cherry.cu)
    dim3 grid((n_tiles + a.tiles_per_cta - 1) / a.tiles_per_cta, n_nodes);
    if (K == 20 && mode == 0) up_wide_kernel<20, 0><<<grid, WT, 0, st>>>(a);
    else if (K == 20) up_wide_kernel<20, 1><<<grid, WT, 0, st>>>(a);
    else if (mode == 0) up_wide_kernel<4, 0><<<grid, WT, 0, st>>>(a);
    else up_wide_kernel<4, 1><<<grid, WT, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_down_wide(int K, const WideArgs &a, int n_tiles, int n_nodes, cudaStream_t st)
{
    if (n_nodes == 0 || n_tiles == 0) return cudaSuccess;
    dim3 grid((n_tiles + a.tiles_per_cta - 1) / a.tiles_per_cta, n_nodes);
    if (K == 20 && a.cherry) down_wide_kernel<20, true><<<grid, WT, 0, st>>>(a);
    else if (K == 20) down_wide_kernel<20, false><<<grid, WT, 0, st>>>(a);
    else down_wide_kernel<4, false><<<grid, WT, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_traceback_wide(int K, const WideArgs &a, int n_nodes, cudaStream_t st)
{
    if (n_nodes == 0) return cudaSuccess;
    dim3 grid((a.ncols + 255) / 256, n_nodes < 512 ? n_nodes : 512);
    traceback_wide_kernel<<<grid, 256, 0, st>>>(a, K, n_nodes);
    return cudaGetLastError();
}

cudaError_t launch_sum_escale(const int16_t *escale, int n_rows, int ncols, int32_t *esum, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(esum, 0, sizeof(int32_t) * ncols, st);
    if (e != cudaSuccess || n_rows == 0) return e;
    dim3 grid((ncols + 127) / 128, (n_rows + 127) / 128);
    sum_escale_kernel<<<grid, 128, 0, st>>>(escale, n_rows, ncols, esum);
    return cudaGetLastError();
}

cudaError_t launch_scatter_rows(const uint8_t *in, uint8_t *out, const int32_t *nodes, int n_rows,
                                const int32_t *orig_col, int ncols, cudaStream_t st)
{
    if (n_rows == 0 || ncols == 0) return cudaSuccess;
    dim3 grid((ncols + 255) / 256, n_rows < 256 ? n_rows : 256);
    scatter_rows_kernel<<<grid, 256, 0, st>>>(in, out, nodes, n_rows, orig_col, ncols);
    return cudaGetLastError();
}

cudaError_t launch_down_wide2(int K, const WideArgs &a, int n_tiles, int n_nodes, cudaStream_t st)
{
    if (n_nodes == 0 || n_tiles == 0) return cudaSuccess;
    if (K != 20 || !a.fuse2) return cudaErrorInvalidValue;
    dim3 grid((n_tiles + a.tiles_per_cta - 1) / a.tiles_per_cta, n_nodes);
    down_wide2_kernel<20><<<grid, WT, 0, st>>>(a);
    return cudaGetLastError();
}

int wide_tile_cols() { return WT; }

}  // namespace bnk
