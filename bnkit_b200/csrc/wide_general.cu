// wide_general.cu -- level kernels (see wide.cu) for the general case: per-column presence masks
// (the pruned forests of Prediction.getTree, src/asr/Prediction.java:329-402) and evidence on any
// node.  Real GRASP alignments are ~60 % gaps, so this is the path real data takes.  One column per
// thread makes the per-column case analysis (absent node, component root, observed node, observed
// parent, ...) plain per-thread control flow; the arithmetic and the order of the log-space additions
// are those of the walker's general kernels (sweeps.cu) and of the oracle:
//   kind NONE  node absent, or present but unconnected (no BN node, PhyloBN.java:149)
//   kind OBS   observed: separates the tree, contributes scalar factors only
//   kind ROOT  unobserved with absent or observed parent: starts from the prior / the observed-parent row
//   kind MSG   unobserved with unobserved present parent: sends a message
// Nodes here have at most two children; with a root-like base the association is
// (base + observed child) + unobserved child, children of equal status in index order.
#include "wide_common.cuh"

namespace bnk {

struct ColInfo {
    bool pv, pp, rootlike, pu[2];
    int ov, op, ou[2], nk, kind;
};

__device__ __forceinline__ ColInfo column_info(const WideArgs &a, const WideNode &w, int ocol)
{
    ColInfo c;
    const size_t nc = a.ncols;
    c.pv = !a.present || a.present[(size_t)w.node * nc + ocol] != 0;
    c.pp = !a.present || a.present[(size_t)w.parent * nc + ocol] != 0;  // wide nodes always have a static parent
    c.ov = a.obs[(size_t)w.node * nc + ocol];
    c.op = c.pp ? a.obs[(size_t)w.parent * nc + ocol] : BNK_NONE;
    c.rootlike = !c.pp || c.op != BNK_NONE;
    c.nk = 0;
#pragma unroll
    for (int e = 0; e < 2; e++) {
        c.pu[e] = e < w.n_child && (!a.present || a.present[(size_t)w.c_node[e] * nc + ocol] != 0);
        c.ou[e] = c.pu[e] ? a.obs[(size_t)w.c_node[e] * nc + ocol] : BNK_NONE;
        c.nk += c.pu[e] ? 1 : 0;
    }
    c.kind = !c.pv ? KIND_NONE : (c.ov != BNK_NONE ? KIND_OBS : ((!c.pp && c.nk == 0) ? KIND_NONE : (c.rootlike ? KIND_ROOT : KIND_MSG)));
    return c;
}

// message of present child e for all K states of this column.  MODE 0: log space, 1: linear.
template <int K, int MODE>
__device__ __forceinline__ void child_vector_gen(const WideArgs &a, const WideNode &w, int e, int ou, size_t tb, int col,
                                                 const double *lt_smem, double (&g)[K])
{
    const int cn = w.c_node[e];
    if (ou != BNK_NONE) {
        if (w.c_ent[e] < 0) {  // childless: its transposed table is staged in shared memory
#pragma unroll
            for (int x = 0; x < K; x++) g[x] = lt_smem[ou * (K + 1) + x];
        } else {               // observed internal node (rare): straight from the table in global memory
            const double *T = a.T_col + (tb + cn) * (K * K) + (size_t)ou * K;
#pragma unroll
            for (int x = 0; x < K; x++) g[x] = T[x];
        }
    } else if (w.c_ent[e] >= 0) {
        const double *m = a.msg + (size_t)w.c_ent[e] * K * a.ncols + col;
#pragma unroll
        for (int x = 0; x < K; x++) g[x] = m[(size_t)x * a.ncols];
    } else {  // uninstantiated childless node: maximise / marginalise its own row
        self_vector<K, MODE>(a.T_row + (tb + cn) * (K * K), g);
    }
}

// ------------------------------------------------------------------------------------
template <int K, int MODE>
__global__ void __launch_bounds__(WT, 4) up_wide_gen_kernel(const WideArgs a)
{
    __shared__ __align__(16) double Ls[K * K];
    __shared__ double LT0[K * (K + 1)], LT1[K * (K + 1)];
    const WideNode w = a.nodes[blockIdx.y];
    const int t_lo = blockIdx.x * a.tiles_per_cta, t_hi = min(t_lo + a.tiles_per_cta, a.n_tiles);
    uint8_t *kind_arr = (MODE == 0) ? a.kind : a.kindm;
    int cur_cat = -1;
    for (int ti = t_lo; ti < t_hi; ti++) {
        const TileDesc tile = a.tiles[ti];
        const size_t tb = (size_t)tile.cat0 * a.n_nodes;
        if (tile.cat0 != cur_cat) {
            __syncthreads();
            load_table<K>(Ls, a.T_row + (tb + w.node) * (K * K), 0);
            if (w.n_child > 0 && w.c_ent[0] < 0) load_table<K>(LT0, a.T_col + (tb + w.c_node[0]) * (K * K), 1);
            if (w.n_child > 1 && w.c_ent[1] < 0) load_table<K>(LT1, a.T_col + (tb + w.c_node[1]) * (K * K), 1);
            tables_wait();
            __syncthreads();
            cur_cat = tile.cat0;
        }
        if ((int)threadIdx.x >= tile.ncols) continue;
        const int col = tile.col0 + threadIdx.x;
        const int ocol = a.orig_col[col];
        const ColInfo c = column_info(a, w, ocol);
        if (kind_arr) kind_arr[(size_t)w.ent * a.ncols + col] = (uint8_t)c.kind;
        if (MODE == 1 && a.escale) a.escale[(size_t)(a.wide_row0 + blockIdx.y) * a.ncols + col] = 0;
        if (MODE == 1 && c.kind == KIND_OBS && a.logl_acc) {  // scalar factors of an observed node (SubstNode.java:582-607)
            double sc = 0.0;
            if (!c.pp) { if (c.nk > 0) sc += log(a.prior[c.ov]); }
            else if (c.op != BNK_NONE) sc += log(Ls[c.op * K + c.ov]);
#pragma unroll
            for (int e = 0; e < 2; e++)  // its observed childless children: P_u[ov][ou]
                if (c.pu[e] && w.c_ent[e] < 0 && c.ou[e] != BNK_NONE) sc += log((e == 0 ? LT0 : LT1)[c.ou[e] * (K + 1) + c.ov]);
            if (sc != 0.0) atomicAdd(a.logl_acc + col, sc);
        }
        if (c.kind != KIND_MSG && c.kind != KIND_ROOT) continue;

        // combine: [base], observed children, unobserved children (each group in index order)
        double G[K], g[K];
        bool have = false;
        if (c.rootlike) {
#pragma unroll
            for (int x = 0; x < K; x++) G[x] = (c.op != BNK_NONE) ? Ls[c.op * K + x] : a.prior[x];
            have = true;
        }
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                if (!c.pu[e] || (c.ou[e] != BNK_NONE) != (pass == 0)) continue;
                child_vector_gen<K, MODE>(a, w, e, c.ou[e], tb, col, e == 0 ? LT0 : LT1, g);
                if (!have) {
#pragma unroll
                    for (int x = 0; x < K; x++) G[x] = g[x];
                    have = true;
                } else {
#pragma unroll
                    for (int x = 0; x < K; x++) G[x] = (MODE == 0) ? G[x] + g[x] : G[x] * g[x];
                }
            }
        }
        if (!have) {
#pragma unroll
            for (int x = 0; x < K; x++) G[x] = (MODE == 0) ? 0.0 : 1.0;
        }
        double *mout = a.msg + (size_t)w.ent * K * a.ncols + col;
        if (MODE == 0) {
            uint8_t *pout = a.ptr + (size_t)w.ent * K * a.ncols + col;
            if (c.kind == KIND_ROOT) {  // first strict maximum of the combined vector; no message
                double best = -CUDART_INF;
                int bi = 0;
#pragma unroll
                for (int x = 0; x < K; x++)
                    if (G[x] > best) { best = G[x]; bi = x; }
                pout[0] = (uint8_t)bi;
                continue;
            }
#pragma unroll 2
            for (int p = 0; p < K; p++) {
                double best;
                int bi;
                argmax_row<K>(Ls + p * K, G, best, bi);
                mout[(size_t)p * a.ncols] = best;
                pout[(size_t)p * a.ncols] = (uint8_t)bi;
            }
        } else {
            int mhi = 0;
#pragma unroll
            for (int x = 0; x < K; x++) mhi = max(mhi, __double2hiint(G[x]));
            int e;
            const double scale = w_pow2_scale(mhi, e);
#pragma unroll
            for (int x = 0; x < K; x++) G[x] *= scale;
            if (a.escale) a.escale[(size_t)(a.wide_row0 + blockIdx.y) * a.ncols + col] = (int16_t)e;
            if (c.kind == KIND_ROOT) {  // a component root: its total goes to the column's log-likelihood
                double tot = 0.0;
#pragma unroll
                for (int x = 0; x < K; x++) tot += G[x];
                if (a.logl_acc) atomicAdd(a.logl_acc + col, log(tot));
                continue;
            }
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
                mout[(size_t)p * a.ncols] = acc0 + acc1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(WT, 3) down_wide_gen_kernel(const WideArgs a)
{
    __shared__ __align__(16) double PTs[K * K];
    __shared__ double LT0[K * (K + 1)], LT1[K * (K + 1)];
    const WideNode w = a.nodes[blockIdx.y];
    const int t_lo = blockIdx.x * a.tiles_per_cta, t_hi = min(t_lo + a.tiles_per_cta, a.n_tiles);
    const int q = a.qrow ? a.qrow[w.ent] : w.ent;
    int cur_cat = -1;
    for (int ti = t_lo; ti < t_hi; ti++) {
        const TileDesc tile = a.tiles[ti];
        const size_t tb = (size_t)tile.cat0 * a.n_nodes;
        if (tile.cat0 != cur_cat) {
            __syncthreads();
            load_table<K>(PTs, a.T_col + (tb + w.node) * (K * K), 0);  // PT[x][xp] = P[xp][x]
            if (w.n_child > 0 && w.c_ent[0] < 0) load_table<K>(LT0, a.T_col + (tb + w.c_node[0]) * (K * K), 1);
            if (w.n_child > 1 && w.c_ent[1] < 0) load_table<K>(LT1, a.T_col + (tb + w.c_node[1]) * (K * K), 1);
            tables_wait();
            __syncthreads();
            cur_cat = tile.cat0;
        }
        if ((int)threadIdx.x >= tile.ncols) continue;
        const int col = tile.col0 + threadIdx.x;
        const int ocol = a.orig_col[col];
        const int kind = a.kindm[(size_t)w.ent * a.ncols + col];
        double *post = (q >= 0 && a.out_post) ? a.out_post + ((size_t)q * a.ncols + ocol) * K : nullptr;
        if (kind != KIND_MSG && kind != KIND_ROOT) {
            if (post) {
#pragma unroll
                for (int x = 0; x < K; x++) post[x] = CUDART_NAN;
            }
            continue;
        }
        const ColInfo c = column_info(a, w, ocol);
        double D[K];
        if (kind == KIND_ROOT) {
#pragma unroll
            for (int x = 0; x < K; x++) D[x] = (c.op != BNK_NONE) ? PTs[x * K + c.op] : a.prior[x];
        } else {
            double T[K];
            const double *t = a.tmsg + (size_t)w.ent * K * a.ncols + col;
            int mhi = 0;
#pragma unroll
            for (int x = 0; x < K; x++) T[x] = t[(size_t)x * a.ncols];
#pragma unroll
            for (int x = 0; x < K; x++) mhi = max(mhi, __double2hiint(T[x]));
            int e;
            const double scale = w_pow2_scale(mhi, e);
#pragma unroll
            for (int x = 0; x < K; x++) T[x] *= scale;
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
        }
        double c0[K], c1[K];
        if (c.pu[0]) child_vector_gen<K, 1>(a, w, 0, c.ou[0], tb, col, LT0, c0);
        else {
#pragma unroll
            for (int x = 0; x < K; x++) c0[x] = 1.0;
        }
        if (c.pu[1]) child_vector_gen<K, 1>(a, w, 1, c.ou[1], tb, col, LT1, c1);
        else {
#pragma unroll
            for (int x = 0; x < K; x++) c1[x] = 1.0;
        }
        const bool push0 = c.pu[0] && c.ou[0] == BNK_NONE && w.c_ent[0] >= 0;
        const bool push1 = c.pu[1] && c.ou[1] == BNK_NONE && w.c_ent[1] >= 0;
        double *t0 = push0 ? a.tmsg + (size_t)w.c_ent[0] * K * a.ncols + col : nullptr;
        double *t1 = push1 ? a.tmsg + (size_t)w.c_ent[1] * K * a.ncols + col : nullptr;
        double S = 0.0;
#pragma unroll
        for (int x = 0; x < K; x++) {
            if (push0) t0[(size_t)x * a.ncols] = D[x] * c1[x];
            if (push1) t1[(size_t)x * a.ncols] = D[x] * c0[x];
            D[x] = D[x] * (c0[x] * c1[x]);
            S += D[x];
        }
        if (post) store_post_row<K>(post, D, w_rcp(S), (reinterpret_cast<uintptr_t>(a.out_post) & 31) == 0);
    }
}

// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) traceback_wide_gen_kernel(const WideArgs a, int K)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col < a.col_lo || col >= a.col_hi) return;
    const WideNode w = a.nodes[blockIdx.y];
    const int kind = a.kind[(size_t)w.ent * a.ncols + col];
    uint8_t s = BNK_NONE;
    if (kind == KIND_MSG) s = a.ptr[((size_t)w.ent * K + a.state_s[(size_t)w.parent * a.ncols + col]) * a.ncols + col];
    else if (kind == KIND_ROOT) s = a.ptr[(size_t)w.ent * K * a.ncols + col];
    else if (kind == KIND_OBS) s = a.obs[(size_t)w.node * a.ncols + a.orig_col[col]];
    a.state_s[(size_t)w.node * a.ncols + col] = s;
}

// ------------------------------------------------------------------------------------
cudaError_t launch_up_wide_gen(int K, int mode, const WideArgs &a, int n_tiles, int n_nodes, cudaStream_t st)
{
    if (n_nodes == 0 || n_tiles == 0) return cudaSuccess;
    dim3 grid((n_tiles + a.tiles_per_cta - 1) / a.tiles_per_cta, n_nodes);
    if (K == 20 && mode == 0) up_wide_gen_kernel<20, 0><<<grid, WT, 0, st>>>(a);
    else if (K == 20) up_wide_gen_kernel<20, 1><<<grid, WT, 0, st>>>(a);
    else if (mode == 0) up_wide_gen_kernel<4, 0><<<grid, WT, 0, st>>>(a);
    else up_wide_gen_kernel<4, 1><<<grid, WT, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_down_wide_gen(int K, const WideArgs &a, int n_tiles, int n_nodes, cudaStream_t st)
{
    if (n_nodes == 0 || n_tiles == 0) return cudaSuccess;
    dim3 grid((n_tiles + a.tiles_per_cta - 1) / a.tiles_per_cta, n_nodes);
    if (K == 20) down_wide_gen_kernel<20><<<grid, WT, 0, st>>>(a);
    else down_wide_gen_kernel<4><<<grid, WT, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_traceback_wide_gen(int K, const WideArgs &a, int n_nodes, cudaStream_t st)
{
    if (n_nodes == 0) return cudaSuccess;
    dim3 grid((a.ncols + 255) / 256, n_nodes);
    traceback_wide_gen_kernel<<<grid, 256, 0, st>>>(a, K);
    return cudaGetLastError();
}

}  // namespace bnk
