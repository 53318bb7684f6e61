// traceback.cu -- max-product traceback (CGTable.getMPE, src/bn/alg/CGTable.java:885-948;
// DenseFactor.getAssign, src/bn/factor/DenseFactor.java:256-276) and small utility kernels.
//
// The reference follows per-factor pointer chains recursively; on a tree that is
//   state[v] = ptr_v[ state[parent(v)] ]     (ptr rows written by the joint up-sweep)
// walked root-to-leaves over the static pre-order program.
#include "device_common.cuh"

namespace bnk {

// One warp owns one column: lane p holds "its" pointer byte ptr_v[p] of each step of a batch
// (the layout the up-sweep wrote; TB_U independent loads in flight per lane), and the chain
//   state[v] = ptr_v[state[parent]]
// is resolved with warp shuffles only: the stack of ancestors' states that are needed later
// (depth bounded by tree_plan.cpp, heavy child last) lives in a register distributed over
// the lanes -- lane k holds slot k -- so one step costs two dependent shuffles and nothing
// else sits on the critical path.  REGSTACK = false keeps the stack in shared memory for
// trees that need more than 32 slots.
constexpr int TB_U = 16;
constexpr int TB_WARPS = 8;

template <bool REGSTACK>
__global__ void __launch_bounds__(TB_WARPS * 32, 5) traceback_kernel(const TracebackArgs a)
{
    extern __shared__ uint8_t tb_smem[];  // [TB_WARPS][tb_slots] when !REGSTACK
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int col = blockIdx.x * TB_WARPS + wid;
    if (col < a.col_lo || col >= a.col_hi) return;
    const int K = a.K;
    uint8_t *st = tb_smem + wid * (a.tb_slots > 0 ? a.tb_slots : 1);
    const uint8_t *pcol = a.ptr + (size_t)col * K + (lane < K ? lane : 0);
    const uint8_t *kcol = a.kind ? a.kind + col : nullptr;
    const unsigned row_stride = (unsigned)(a.ncols * K), ncols_u = (unsigned)a.ncols;
    const unsigned FULL = 0xffffffffu;
    int stack_reg = BNK_NONE;  // REGSTACK: slot `lane`

    // lane u of the warp carries the program fields of step b0 + u
    int n_ent = 0, n_node = 0, n_meta = 0;
    auto fetch = [&](int b0) {
        const int si = b0 + lane;
        if (lane < TB_U && si < a.n_steps) {
            const Step &sp = a.down[si];
            n_ent = sp.ent; n_node = sp.node;
            n_meta = (sp.pslot + 1) | ((sp.oslot + 1) << 12) | ((sp.parent < 0 ? 1 : 0) << 24);
        }
    };
    fetch(0);
    for (int b0 = 0; b0 < a.n_steps; b0 += TB_U) {
        const int c_ent = n_ent, c_node = n_node, c_meta = n_meta;
        fetch(b0 + TB_U);
        int byte[TB_U];
#pragma unroll
        for (int u = 0; u < TB_U; u++) {
            const unsigned ent = (unsigned)__shfl_sync(FULL, c_ent, u);
            byte[u] = (b0 + u < a.n_steps) ? pcol[(size_t)ent * row_stride] : 0;
        }
#pragma unroll
        for (int u = 0; u < TB_U; u++) {
            if (b0 + u >= a.n_steps) break;
            const int node = __shfl_sync(FULL, c_node, u), meta = __shfl_sync(FULL, c_meta, u);
            const int pslot = (meta & 0xfff) - 1, oslot = ((meta >> 12) & 0xfff) - 1;
            int kind = (meta >> 24) ? KIND_ROOT : KIND_MSG;
            if (kcol) kind = kcol[(size_t)(unsigned)__shfl_sync(FULL, c_ent, u) * ncols_u];
            int src = 0;
            if (REGSTACK) {
                const int ps = __shfl_sync(FULL, stack_reg, pslot & 31);
                if (kind == KIND_MSG) src = ps;
            } else if (kind == KIND_MSG) {
                src = st[pslot];
            }
            int s = __shfl_sync(FULL, byte[u], src & 31);
            if (kind == KIND_OBS) s = a.obs[(size_t)(unsigned)node * ncols_u + a.orig_col[col]];
            else if (kind != KIND_MSG && kind != KIND_ROOT) s = BNK_NONE;
            if (REGSTACK) {
                if (lane == oslot) stack_reg = s;
            } else {
                if (lane == 0 && oslot >= 0) st[oslot] = (uint8_t)s;
                __syncwarp();
            }
            if (lane == 0) a.state_s[(size_t)(unsigned)node * ncols_u + col] = (uint8_t)s;
        }
    }
}

cudaError_t launch_traceback(const TracebackArgs &a, cudaStream_t st)
{
    if (a.tb_slots >= 4095) return cudaErrorInvalidValue;
    const unsigned grid = (a.ncols + TB_WARPS - 1) / TB_WARPS;
    if (a.tb_slots <= 32) {
        traceback_kernel<true><<<grid, TB_WARPS * 32, 0, st>>>(a);
    } else {
        const size_t smem = (size_t)TB_WARPS * a.tb_slots;
        traceback_kernel<false><<<grid, TB_WARPS * 32, smem, st>>>(a);
    }
    return cudaGetLastError();
}

// Childless nodes: the observed state passes through (MaxLhoodJoint.java:161-163); an
// uninstantiated childless node under a present parent gets argmax_x L_u[state(parent)][x]
// (first strict maximum), or the root-like rule when the parent is observed.
__global__ void leaf_states_kernel(const TracebackArgs a)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= a.ncols) return;
    const int gcat = a.cat_of_col[col];
    if (gcat < a.cat_lo || gcat >= a.cat_hi) return;
    for (int li = blockIdx.y; li < a.n_leaves; li += gridDim.y) {
        const int u = a.leaf_nodes[li], par = a.leaf_parent[li];
        const size_t uc = (size_t)u * a.ncols + col;
        uint8_t s = BNK_NONE;
        if (a.present == nullptr || a.present[uc]) {
            s = a.obs[uc];
            if (s == BNK_NONE && par >= 0 && (a.present == nullptr || a.present[(size_t)par * a.ncols + col])) {
                // parent's state: observed, or already written by the traceback kernels
                const uint8_t ps = a.out_state[(size_t)par * a.ncols + col];
                if (ps != BNK_NONE) {
                    const int K = a.K;
                    const int cat = gcat - a.cat_lo;
                    const double *row = a.L + (((size_t)cat * a.n_nodes + u) * K + ps) * K;
                    double best = -CUDART_INF;
                    int bi = 0;
                    for (int x = 0; x < K; x++)
                        if (row[x] > best) { best = row[x]; bi = x; }
                    s = (uint8_t)bi;
                }
            }
        }
        a.out_state[uc] = s;
    }
}

cudaError_t launch_leaf_states(const TracebackArgs &a, cudaStream_t st)
{
    if (a.n_leaves == 0 || a.ncols == 0) return cudaSuccess;
    dim3 grid((a.ncols + 255) / 256, a.n_leaves < 8192 ? a.n_leaves : 8192);
    leaf_states_kernel<<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

// Posterior of an uninstantiated CHILDLESS node v (possible after orphan flooding, IdxTree.java:429-432): the
// reference's elimination gives it a distribution like any other variable.  v sends its parent p the message
// g_v[xp] = sum_x P_v[xp][x], so with the parent's (normalised) posterior U_p the vector from above is
// T[xp] = U_p[xp] / g_v[xp] and post_v[x] ~ sum_xp P_v[xp][x] T[xp]; under an observed parent it is the
// normalised row P_v[obs_p][.].  One thread per column; rows of NaN where v has no BN variable.
__global__ void leaf_post_kernel(const LeafPostArgs a)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;  // ORIGINAL column
    if (col >= a.ncols) return;
    const int gcat = a.cat_of_col[col];
    if (gcat < a.cat_lo || gcat >= a.cat_hi) return;
    const int K = a.K;
    double *out = a.out_row + (size_t)col * K;
    const size_t vc = (size_t)a.node * a.ncols + col, pc = (size_t)(a.parent < 0 ? 0 : a.parent) * a.ncols + col;
    const bool pv = a.present == nullptr || a.present[vc];
    const bool pp = a.parent >= 0 && (a.present == nullptr || a.present[pc]);
    bool ok = pv && pp && a.obs[vc] == BNK_NONE;
    const double *P = a.P + ((size_t)(gcat - a.cat_lo) * a.n_nodes + a.node) * K * K;
    double y[BNK_MAXK];
    if (ok) {
        const int op = a.obs[pc];
        if (op != BNK_NONE) {
            for (int x = 0; x < K; x++) y[x] = P[op * K + x];
        } else {
            const double *up = a.parent_row + (size_t)col * K;
            if (up[0] != up[0]) ok = false;
            else {
                for (int x = 0; x < K; x++) y[x] = 0.0;
                for (int xp = 0; xp < K; xp++) {
                    double g = 0.0;
                    for (int x = 0; x < K; x++) g += P[xp * K + x];
                    const double t = up[xp] / g;
                    for (int x = 0; x < K; x++) y[x] = fma(P[xp * K + x], t, y[x]);
                }
            }
        }
    }
    if (ok) {
        double s = 0.0;
        for (int x = 0; x < K; x++) s += y[x];
        for (int x = 0; x < K; x++) out[x] = y[x] / s;
    } else {
        for (int x = 0; x < K; x++) out[x] = CUDART_NAN;
    }
}

cudaError_t launch_leaf_post(const LeafPostArgs &a, cudaStream_t st)
{
    if (a.ncols == 0) return cudaSuccess;
    leaf_post_kernel<<<(a.ncols + 127) / 128, 128, 0, st>>>(a);
    return cudaGetLastError();
}

// out[row][c] = in[row][orig_col[c]]
__global__ void gather_cols_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                                   const int32_t *__restrict__ orig_col, int n_rows, int ncols)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    const int oc = orig_col[c];
    for (int r = blockIdx.y; r < n_rows; r += gridDim.y) out[(size_t)r * ncols + c] = in[(size_t)r * ncols + oc];
}

cudaError_t launch_gather_cols(const uint8_t *in, uint8_t *out, const int32_t *orig_col, int n_rows,
                               int ncols, cudaStream_t st)
{
    if (n_rows == 0 || ncols == 0) return cudaSuccess;
    dim3 grid((ncols + 255) / 256, n_rows < 4096 ? n_rows : 4096);
    gather_cols_kernel<<<grid, 256, 0, st>>>(in, out, orig_col, n_rows, ncols);
    return cudaGetLastError();
}

// One pass over the whole observation matrix: flag |= 1 if a node with children carries an observation (the
// lean kernels assume evidence on childless nodes only), flag |= 2 if any byte is neither BNK_NONE nor a state
// index < K (it would index past the K x K tables: ADVICE r1).  Rows are read 4 bytes per thread when aligned.
__global__ void __launch_bounds__(256) scan_obs_kernel(const uint8_t *__restrict__ obs, const uint8_t *__restrict__ has_child,
                                                       int n_nodes, int ncols, int K, int vec4, int32_t *flag)
{
    int found = 0;
    if (vec4) {
        const int c4 = blockIdx.x * blockDim.x + threadIdx.x;
        if (c4 * 4 < ncols)
            for (int v = blockIdx.y; v < n_nodes; v += gridDim.y) {
                const uint32_t w = *reinterpret_cast<const uint32_t *>(obs + (size_t)v * ncols + 4 * c4);
                if (w == 0xffffffffu) continue;
                const int hc = has_child[v];
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int o = (w >> (8 * b)) & 0xff;
                    if (o != BNK_NONE) found |= (hc ? 1 : 0) | (o >= K ? 2 : 0);
                }
            }
    } else {
        const int c = blockIdx.x * blockDim.x + threadIdx.x;
        if (c < ncols)
            for (int v = blockIdx.y; v < n_nodes; v += gridDim.y) {
                const int o = obs[(size_t)v * ncols + c];
                if (o != BNK_NONE) found |= (has_child[v] ? 1 : 0) | (o >= K ? 2 : 0);
            }
    }
    const int any1 = __syncthreads_or(found & 1), any2 = __syncthreads_or(found & 2);
    if (threadIdx.x == 0 && (any1 || any2)) atomicOr(flag, (any1 ? 1 : 0) | (any2 ? 2 : 0));
}

cudaError_t launch_scan_obs(const uint8_t *obs, const uint8_t *has_child, int n_nodes, int ncols, int K,
                            int32_t *flag, cudaStream_t st)
{
    if (n_nodes == 0 || ncols == 0) return cudaSuccess;
    const int vec4 = (ncols % 4 == 0) && ((reinterpret_cast<uintptr_t>(obs) & 3) == 0);
    const int per = vec4 ? ncols / 4 : ncols;
    dim3 grid((per + 255) / 256, n_nodes < 4096 ? n_nodes : 4096);
    scan_obs_kernel<<<grid, 256, 0, st>>>(obs, has_child, n_nodes, ncols, K, vec4, flag);
    return cudaGetLastError();
}

__global__ void fill_f64_kernel(double *p, size_t n, double v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
cudaError_t launch_fill_f64(double *p, size_t n, double v, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    fill_f64_kernel<<<blocks, 256, 0, st>>>(p, n, v);
    return cudaGetLastError();
}

// deterministic single-CTA sum (column log-likelihoods -> one double for the all-reduce)
__global__ void __launch_bounds__(1024) sum_f64_kernel(const double *__restrict__ in, int n, double *out)
{
    __shared__ double part[1024];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += in[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) part[threadIdx.x] += part[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = part[0];
}
cudaError_t launch_sum_f64(const double *in, int n, double *out, cudaStream_t st)
{
    sum_f64_kernel<<<1, 1024, 0, st>>>(in, n, out);
    return cudaGetLastError();
}

}  // namespace bnk
