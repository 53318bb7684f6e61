// traceback.cu -- max-product traceback (CGTable.getMPE, src/bn/alg/CGTable.java:885-948;
// DenseFactor.getAssign, src/bn/factor/DenseFactor.java:256-276) and small utility kernels.
//
// The reference follows per-factor pointer chains recursively; on a tree that is
//   state[v] = ptr_v[ state[parent(v)] ]     (ptr rows written by the joint up-sweep)
// walked root-to-leaves.  One thread owns one column and walks the static pre-order
// program; ancestors' states needed later sit in a per-thread shared-memory stack whose
// depth is bounded by tree_plan.cpp (heavy child last).
#include "device_common.cuh"

namespace bnk {

__global__ void __launch_bounds__(128) traceback_kernel(const TracebackArgs a)
{
    extern __shared__ uint8_t sstack[];  // [tb_slots][blockDim.x]
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= a.ncols) return;
    const int gcat = a.cat_of_col[col];
    if (gcat < a.cat_lo || gcat >= a.cat_hi) return;
    const int K = a.K, nt = blockDim.x, tid = threadIdx.x;
    for (int i = 0; i < a.n_steps; i++) {
        const DownStep st = a.down[i];
        const size_t ec = (size_t)st.ent * a.ncols + col;
        const uint8_t kind = a.kind[ec];
        uint8_t s = BNK_NONE;
        if (kind == KIND_MSG) {
            const uint8_t ps = sstack[st.pslot * nt + tid];
            s = a.ptr[ec * K + ps];
        } else if (kind == KIND_ROOT) {
            s = a.ptr[ec * K];
        } else if (kind == KIND_OBS) {
            s = a.obs[(size_t)st.node * a.ncols + col];
        }
        if (st.oslot >= 0) sstack[st.oslot * nt + tid] = s;
        a.out_state[(size_t)st.node * a.ncols + col] = s;
    }
}

cudaError_t launch_traceback(const TracebackArgs &a, cudaStream_t st)
{
    const int threads = 128;
    const size_t smem = (size_t)(a.tb_slots > 0 ? a.tb_slots : 1) * threads;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(traceback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    traceback_kernel<<<(a.ncols + threads - 1) / threads, threads, smem, st>>>(a);
    return cudaGetLastError();
}

// Childless nodes: the observed state passes through (MaxLhoodJoint.java:161-163); an
// uninstantiated childless node under a present parent gets argmax_x L_u[state(parent)][x]
// (first strict maximum), or the root-like rule when the parent is observed.
__global__ void leaf_states_kernel(const TracebackArgs a)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)a.n_leaves * a.ncols) return;
    const int li = (int)(idx / a.ncols), col = (int)(idx % a.ncols);
    const int gcat = a.cat_of_col[col];
    if (gcat < a.cat_lo || gcat >= a.cat_hi) return;
    const int u = a.leaf_nodes[li], par = a.leaf_parent[li];
    const size_t uc = (size_t)u * a.ncols + col;
    uint8_t s = BNK_NONE;
    if (a.present == nullptr || a.present[uc]) {
        s = a.obs[uc];
        if (s == BNK_NONE && par >= 0 && (a.present == nullptr || a.present[(size_t)par * a.ncols + col])) {
            // parent's state: observed, or already written by the traceback kernel
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

cudaError_t launch_leaf_states(const TracebackArgs &a, cudaStream_t st)
{
    const size_t total = (size_t)a.n_leaves * a.ncols;
    if (total == 0) return cudaSuccess;
    leaf_states_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
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

__global__ void scan_internal_obs_kernel(const uint8_t *__restrict__ obs, const int32_t *__restrict__ node_of_ent,
                                         int n_int, int ncols, int32_t *flag)
{
    const size_t total = (size_t)n_int * ncols;
    int found = 0;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(idx / ncols), c = (int)(idx % ncols);
        if (obs[(size_t)node_of_ent[e] * ncols + c] != BNK_NONE) found = 1;
    }
    if (__syncthreads_or(found) && threadIdx.x == 0) atomicOr(flag, 1);
}

cudaError_t launch_scan_internal_obs(const uint8_t *obs, const int32_t *node_of_ent, int n_int, int ncols,
                                     int32_t *flag, cudaStream_t st)
{
    if (n_int == 0 || ncols == 0) return cudaSuccess;
    scan_internal_obs_kernel<<<148 * 4, 256, 0, st>>>(obs, node_of_ent, n_int, ncols, flag);
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
