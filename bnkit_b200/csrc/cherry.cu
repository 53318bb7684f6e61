// cherry.cu -- up-sweeps of the height-1 level ("cherries": every child is a childless node) with
// site-pattern compression.
//
// The message of a cherry depends on the column only through the observed states of its (at most two)
// children and the column's rate category: M_v = f(o_0, o_1).  Half of the internal nodes of a binary tree
// sit at height 1, and among the ~1,000 columns of one category only a few percent to ~20 % of the
// (o_0, o_1) pairs are distinct (sister leaves are similar), so a CTA that owns (node, category, <= CG
// columns) lists the distinct pairs, evaluates the 400-candidate max-out / 400-FMA sum-out once per pair
// (one pair per thread, the arithmetic and its order exactly those of up_wide_kernel, wide.cu) into a
// per-(node, group) record table in global memory, and then streams the result rows out to every column -- the store-all-partials traffic of SURVEY.md 8(d) is
// unchanged, the issue-bound arithmetic shrinks by the repeat factor.  Bit-identical to the level kernel.
// Preconditions (host): lean path (nothing masked, evidence on childless nodes only), <= 2 children.
#include "wide_common.cuh"

namespace bnk {

template <int K, int MODE>
__global__ void __launch_bounds__(WT, 5) cherry_pairs_kernel(const WideArgs a)
{
    constexpr int NS = K + 1;  // child codes: state 0..K-1, K = uninstantiated
    constexpr int NCODE = NS * NS;
    constexpr int CPT = (NCODE + WT - 1) / WT;  // codes per thread during compaction
    __shared__ __align__(16) double Ls[K * K];
    __shared__ double LT0[K * (K + 1)], LT1[K * (K + 1)];
    __shared__ int16_t slot[NCODE], list[NCODE], ccode[CG], pexp[NCODE];
    __shared__ int wsum[WT / 32];

    const int tid = threadIdx.x;
    const WideNode w = a.nodes[blockIdx.y];
    const TileDesc g = a.cgroups[blockIdx.x];
    const size_t tb = (size_t)g.cat0 * a.n_nodes;
    const bool two = w.n_child > 1;
    const size_t row = (size_t)a.wide_row0 + blockIdx.y;  // the node's row in the height-1 level (the level starts the wide list)
    const CherryTable<K> tab(a.cscratch, row, a.crecs_per_node, g);
    load_table<K>(Ls, a.T_row + (tb + w.node) * (K * K), 0);
    load_table<K>(LT0, a.T_col + (tb + w.c_node[0]) * (K * K), 1);
    if (two) load_table<K>(LT1, a.T_col + (tb + w.c_node[1]) * (K * K), 1);
    for (int q = tid; q < NCODE; q += WT) slot[q] = 0;
    __syncthreads();  // (the tables are still in flight: they are waited for after the state loads below)

    // 1. the pair code of every column; mark the codes that occur.  All index loads are issued before the
    //    first state load so the group costs two memory round trips.
    constexpr int NQ = CG / WT;
    {
        const uint8_t *o0p = a.obs + (size_t)w.c_node[0] * a.ncols;
        const uint8_t *o1p = a.obs + (size_t)w.c_node[two ? 1 : 0] * a.ncols;
        int oc[NQ];
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const int i = tid + q * WT;
            oc[q] = i < g.ncols ? a.orig_col[g.col0 + i] : -1;  // inputs stay in the caller's column order
        }
        int o0[NQ], o1[NQ];
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            o0[q] = oc[q] >= 0 ? o0p[oc[q]] : 0;
            o1[q] = (two && oc[q] >= 0) ? o1p[oc[q]] : 0;
        }
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const int c0 = o0[q] >= K ? K : o0[q], c1 = o1[q] >= K ? K : o1[q];
            if (oc[q] >= 0) { ccode[tid + q * WT] = (int16_t)(c0 * NS + c1); slot[c0 * NS + c1] = 1; }
        }
    }
    tables_wait();
    __syncthreads();

    // 2. compact: slot[code] = rank among the occurring codes, list[rank] = code
    int npairs;
    {
        int cnt = 0, flags = 0;
#pragma unroll
        for (int q = 0; q < CPT; q++) {
            const int c = tid * CPT + q;
            if (c < NCODE && slot[c]) { flags |= 1 << q; cnt++; }
        }
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d *= 2) {
            const int n = __shfl_up_sync(0xffffffffu, incl, d);
            if ((tid & 31) >= d) incl += n;
        }
        if ((tid & 31) == 31) wsum[tid >> 5] = incl;
        __syncthreads();
        int base = 0;
        npairs = 0;
#pragma unroll
        for (int q = 0; q < WT / 32; q++) {
            if (q < (tid >> 5)) base += wsum[q];
            npairs += wsum[q];
        }
        int r = base + incl - cnt;
#pragma unroll
        for (int q = 0; q < CPT; q++)
            if (flags >> q & 1) {
                const int c = tid * CPT + q;
                slot[c] = (int16_t)r;
                list[r] = (int16_t)c;
                r++;
            }
    }
    __syncthreads();

    // 3. one pair per thread: its record goes to the scratch area
    for (int j = tid; j < npairs; j += WT) {
        const int code = list[j];
        const int o0 = code / NS, o1 = code - o0 * NS;
        double G[K];
        if (o0 < K) {
#pragma unroll
            for (int x = 0; x < K; x++) G[x] = LT0[o0 * (K + 1) + x];
        } else {
            self_vector<K, MODE>(a.T_row + (tb + w.c_node[0]) * (K * K), G);
        }
        if (two) {
            double gg[K];
            if (o1 < K) {
#pragma unroll
                for (int x = 0; x < K; x++) gg[x] = LT1[o1 * (K + 1) + x];
            } else {
                self_vector<K, MODE>(a.T_row + (tb + w.c_node[1]) * (K * K), gg);
            }
#pragma unroll
            for (int x = 0; x < K; x++) G[x] = (MODE == 0) ? G[x] + gg[x] : G[x] * gg[x];
        }
        if (MODE == 0) {
#pragma unroll 2
            for (int p = 0; p < K; p++) {
                double best;
                int bi;
                argmax_row<K>(Ls + p * K, G, best, bi);
                tab.v[p * tab.cap + j] = best;
                tab.ptr[p * tab.cap + j] = (uint8_t)bi;
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
                tab.v[p * tab.cap + j] = acc0 + acc1;
            }
            tab.e[j] = e;
            pexp[j] = (int16_t)e;
        }
    }
    // the record index of every column (for whoever reads the records: the expansion kernel, or the parent level
    // directly) and, summing into the log-likelihood later, the column's rescale exponent
    int16_t *rank_g = a.cslots + row * a.ncols + g.col0;
    for (int i = tid; i < g.ncols; i += WT) rank_g[i] = slot[ccode[i]];
    if (MODE == 1 && a.escale) {
        __syncthreads();
        int16_t *es = a.escale + row * a.ncols + g.col0;
        for (int i = tid; i < g.ncols; i += WT) es[i] = pexp[slot[ccode[i]]];
    }
}

// Expansion: one (node, column) per thread, the same warp-row store pattern as the level kernels.  Each thread
// reads its column's record index, and copies the record's K
// values (and pointers / exponent) into the message rows.  Pure streaming: no shared memory, full occupancy.
// Tiles (a.ctiles) start on 16-column boundaries of the absolute sorted column index, so every warp writes full,
// sector-aligned 256-byte row segments (tools/microbench/write_bw.cu).
template <int K, int MODE>
__global__ void __launch_bounds__(WT, 8) cherry_expand_kernel(const WideArgs a)
{
    const TileDesc t = a.ctiles[blockIdx.x];  // col0 (aligned start), ncols (= WT), cat0 = group
    const TileDesc g = a.cgroups[t.cat0];
    const int c = t.col0 + threadIdx.x;
    if (c < g.col0 || c >= g.col0 + g.ncols) return;
    const WideNode w = a.nodes[blockIdx.y];
    // a.rest_rows: the expansion runs on a subset of the level (nodes under a walked parent), rows given explicitly
    const size_t row = a.rest_rows ? (size_t)a.rest_rows[blockIdx.y] : (size_t)a.wide_row0 + blockIdx.y;
    const CherryTable<K> tab(a.cscratch, row, a.crecs_per_node, g);
    const int r = a.cslots[row * a.ncols + c];
    double v[K];
#pragma unroll
    for (int p = 0; p < K; p++) v[p] = __ldg(tab.v + p * tab.cap + r);
    uint8_t pk[MODE == 0 ? K : 1];
    if (MODE == 0) {
#pragma unroll
        for (int p = 0; p < K; p++) pk[p] = __ldg(tab.ptr + p * tab.cap + r);
    }
    double *m = a.msg + (size_t)w.ent * K * a.ncols + c;
#pragma unroll
    for (int p = 0; p < K; p++) m[(size_t)p * a.ncols] = v[p];
    if (MODE == 0) {
        uint8_t *pp = a.ptr + (size_t)w.ent * K * a.ncols + c;
#pragma unroll
        for (int p = 0; p < K; p++) pp[(size_t)p * a.ncols] = pk[p];
    }
}

size_t cherry_record_bytes(int K) { return K == 20 ? CherryTable<20>::REC_BYTES : CherryTable<4>::REC_BYTES; }
int cherry_group_capacity(int K, int ncols) { return cherry_cap(K, ncols); }

// the two halves separately: cherry-direct runs the pairs kernel on the whole level and expands only the nodes
// under a walked parent (a.nodes = that list, a.rest_rows = their rows in the level)
cudaError_t launch_cherry_pairs(int K, int mode, const WideArgs &a, int n_nodes, cudaStream_t st)
{
    if (n_nodes == 0 || a.n_cgroups == 0) return cudaSuccess;
    dim3 grid(a.n_cgroups, n_nodes);
    if (K == 20 && mode == 0) cherry_pairs_kernel<20, 0><<<grid, WT, 0, st>>>(a);
    else if (K == 20) cherry_pairs_kernel<20, 1><<<grid, WT, 0, st>>>(a);
    else if (mode == 0) cherry_pairs_kernel<4, 0><<<grid, WT, 0, st>>>(a);
    else cherry_pairs_kernel<4, 1><<<grid, WT, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_cherry_expand(int K, int mode, const WideArgs &a, int n_nodes, cudaStream_t st)
{
    if (n_nodes == 0 || a.n_cgroups == 0) return cudaSuccess;
    dim3 grid2(a.n_ctiles, n_nodes);
    if (K == 20 && mode == 0) cherry_expand_kernel<20, 0><<<grid2, WT, 0, st>>>(a);
    else if (K == 20) cherry_expand_kernel<20, 1><<<grid2, WT, 0, st>>>(a);
    else if (mode == 0) cherry_expand_kernel<4, 0><<<grid2, WT, 0, st>>>(a);
    else cherry_expand_kernel<4, 1><<<grid2, WT, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_up_cherry(int K, int mode, const WideArgs &a, int n_nodes, cudaStream_t st)
{
    cudaError_t e = launch_cherry_pairs(K, mode, a, n_nodes, st);
    if (e != cudaSuccess) return e;
    return launch_cherry_expand(K, mode, a, n_nodes, st);
}

// Traceback of the height-1 level straight from the record tables (cherry-direct): state = ptr record of the
// column's pair, entry state(parent).  One (node, column) per thread; the group of a column by a scan of the
// (few) groups.
template <int K>
__global__ void __launch_bounds__(256) traceback_cherry_kernel(const WideArgs a)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col < a.col_lo || col >= a.col_hi) return;
    const WideNode w = a.nodes[blockIdx.y];
    const size_t row = (size_t)a.wide_row0 + blockIdx.y;
    int gi = 0;
    while (gi + 1 < a.n_cgroups && col >= a.cgroups[gi].col0 + a.cgroups[gi].ncols) gi++;
    const TileDesc g = a.cgroups[gi];
    const CherryTable<K> tab(a.cscratch, row, a.crecs_per_node, g);
    const int r = a.cslots[row * a.ncols + col];
    const int sp = a.state_s[(size_t)w.parent * a.ncols + col];
    a.state_s[(size_t)w.node * a.ncols + col] = tab.ptr[(size_t)sp * tab.cap + r];
}
cudaError_t launch_traceback_cherry(int K, const WideArgs &a, int n_nodes, cudaStream_t st)
{
    if (n_nodes == 0) return cudaSuccess;
    dim3 grid((a.ncols + 255) / 256, n_nodes);
    if (K == 20) traceback_cherry_kernel<20><<<grid, 256, 0, st>>>(a);
    else traceback_cherry_kernel<4><<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

int cherry_group_cols() { return CG; }

}  // namespace bnk
