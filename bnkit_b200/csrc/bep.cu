// bep.cu -- bi-directional edge parsimony (BEP), the reference's default indel inference: which ancestors
// exist in which alignment column.  It is the step right before the substitution-model sweeps:
// its result is the per-column presence mask those sweeps take.
//
// Reference: asr.Prediction.PredictByBidirEdgeParsimony (src/asr/Prediction.java:1443-1524) builds, for every
// column i in -1..nPos and both directions, a TreeInstance whose leaf values are "the next (previous) occupied
// column of this sequence" (POGTree.getEdgeInstance, src/dat/pog/POGTree.java:450-487) and runs
// asr.Parsimony on it (unit-cost Sankoff keeping ALL optimal states, src/asr/Parsimony.java:215-412), one job
// per instance on a thread pool (ThreadedDecorators).  The instances are independent: here one THREAD owns one
// instance and the whole tree, lanes of a warp own neighbouring columns, so every load and store of the
// per-node state is a coalesced row.
//
// Unit costs make the Sankoff tables collapse to bit sets.  With m_c = min_j S_c(j) a child contributes
// min(S_c(i), m_c + 1) = m_c + [i not in A_c] to its parent's score for state i, where A_c is the child's arg-min
// set, so S_b(i) = sum_c m_c + #{c : i not in A_c}: only A (score == min) and B (score == min + 1) are ever
// needed.  Forward: count, per state, the children whose A misses it (bit-sliced counters), A_b / B_b = states at
// the lowest / next count.  Backward (all optimal states, Parsimony.java:327-389): opt_root = A_root and
//   opt_c = (opt_b & (A_c | B_c)) | (A_c if opt_b has a state outside A_c)
// which is exactly the union over parent states i of the child states j minimising S_c(j) + [i != j].
// oracle/bep.py keeps the reference's explicit score tables and traceback lists; tests compare the two.
#include "bep.cuh"

namespace bnk {


// next / prev occupied column of every childless node, extended index e = column + 1 in 0 .. n_pos + 1:
//   nxt[leaf][e] (e <= n_pos)  = POGTree.getEdgeInstance(e - 1, forward)  for that leaf
//   prv[leaf][e] (e >= 1)      = POGTree.getEdgeInstance(e - 1, backward)
__global__ void bep_next_prev_kernel(const uint8_t *obs, const int32_t *leaf_nodes, int n_leaves, int n_pos, int32_t *nxt,
                                     int32_t *prv)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_leaves) return;
    const uint8_t *row = obs + (size_t)leaf_nodes[l] * n_pos;
    int32_t *nx = nxt + (size_t)l * (n_pos + 2), *pv = prv + (size_t)l * (n_pos + 2);
    int cur = BEP_NONE;  // next occupied column seen so far, scanning right to left
    for (int c = n_pos - 1; c >= 0; c--) {
        const bool occ = row[c] != BNK_NONE;
        nx[c + 1] = occ ? (cur == BEP_NONE ? n_pos : cur) : BEP_NONE;
        if (occ) cur = c;
    }
    nx[0] = cur;               // column -1: the first occupied column (none: the sequence is empty)
    nx[n_pos + 1] = BEP_NONE;  // (nPos, forward) is never instantiated
    cur = BEP_NONE;
    for (int c = 0; c < n_pos; c++) {
        const bool occ = row[c] != BNK_NONE;
        pv[c + 1] = occ ? (cur == BEP_NONE ? -1 : cur) : BEP_NONE;
        if (occ) cur = c;
    }
    pv[n_pos + 1] = cur;
    pv[0] = BEP_NONE;
}


__device__ __forceinline__ int bep_leaf_value(const BepArgs &a, int leaf, int e, bool fwd)
{
    return (fwd ? a.nxt : a.prv)[(size_t)leaf * (a.n_pos + 2) + e];
}

// Symbols of an instance = its distinct non-null leaf values (TreeInstance ctor; the order is immaterial to the
// bit-set sweeps, so values are numbered in ascending order).  Two passes over a byte map [instance][value + 1]:
// every (leaf, instance) pair marks its value, then one warp per instance numbers the marked values (ballot +
// prefix count), leaving symbol + 1 in the map -- the O(1) lookup the sweeps use for childless nodes -- and the
// value of every symbol in dict[symbol][instance] for the host.
__global__ void bep_mark_kernel(const BepArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_inst) return;
    const int e = a.inst_e[t];
    const bool fwd = a.inst_fwd[t] != 0;
    uint8_t *row = a.symmap + (size_t)t * (a.n_pos + 2);
    for (int l = blockIdx.y; l < a.n_leaves; l += gridDim.y) {
        const int v = bep_leaf_value(a, l, e, fwd);
        if (v != BEP_NONE) row[v + 1] = 1;
    }
}
__global__ void bep_number_kernel(const BepArgs a)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= a.n_inst) return;
    uint8_t *row = a.symmap + (size_t)t * (a.n_pos + 2);
    int n = 0;
    for (int base = 0; base < a.n_pos + 2; base += 32) {
        const int i = base + lane;
        const bool on = i < a.n_pos + 2 && row[i] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if (on) {
            const int k = n + __popc(m & ((1u << lane) - 1u));
            if (k < a.max_sym) { row[i] = (uint8_t)(k + 1); a.dict[(size_t)k * a.n_inst + t] = i - 1; }
            else row[i] = 0;
        }
        n += __popc(m);
    }
    if (lane == 0) {
        if (n > a.max_sym) { atomicOr(a.flag, 1); n = a.max_sym; }
        a.nsym[t] = n;
    }
}

// symbol bit of a childless node: its value's symbol, the NULL symbol, or "every state costs nothing"
template <int W>
__device__ __forceinline__ void bep_leaf_set(const BepArgs &a, int t, int leaf, int e, bool fwd, int n, const uint32_t *valid,
                                             uint32_t *out)
{
#pragma unroll
    for (int w = 0; w < W; w++) out[w] = 0;
    const int v = bep_leaf_value(a, leaf, e, fwd);
    if (v == BEP_NONE) {
        if (a.recode_null) out[n >> 5] |= 1u << (n & 31);  // NULL = symbol n
        else {
#pragma unroll
            for (int w = 0; w < W; w++) out[w] = valid[w];  // un-instantiated leaf (Parsimony.java:259-264)
        }
        return;
    }
    const int k = (int)a.symmap[(size_t)t * (a.n_pos + 2) + v + 1] - 1;
    if (k >= 0) out[k >> 5] |= 1u << (k & 31);
}

template <int W>
__device__ __forceinline__ void bep_valid_mask(const BepArgs &a, int n, uint32_t *valid)
{
    const int ntot = n + (a.recode_null ? 1 : 0);
#pragma unroll
    for (int w = 0; w < W; w++) {
        const int lo = w * 32;
        valid[w] = ntot >= lo + 32 ? 0xffffffffu : (ntot > lo ? (1u << (ntot - lo)) - 1u : 0u);
    }
}

// forward step of one internal node v (row ent) of instance t: A = states at the lowest count of children whose
// arg-min set misses them, B = states one count above
template <int W, int P>   // P = bits of the bit-sliced counters: up to 2^P - 1 children per node
__device__ __forceinline__ void bep_node_forward(const BepArgs &a, int t, int v, int ent, int e, bool fwd, int n, const uint32_t *valid)
{
    const size_t ni = (size_t)a.n_inst;
    uint32_t plane[P][W];
#pragma unroll
    for (int p = 0; p < P; p++)
#pragma unroll
        for (int w = 0; w < W; w++) plane[p][w] = 0;
    const int c0 = a.first[v], c1 = a.first[v + 1];
    for (int ci = c0; ci < c1; ci++) {
        const int u = a.kids[ci];
        const int ue = a.ent_of_node[u];
        uint32_t Ac[W];
        if (ue >= 0) {
#pragma unroll
            for (int w = 0; w < W; w++) Ac[w] = a.A[((size_t)ue * W + w) * ni + t];
        } else {
            bep_leaf_set<W>(a, t, a.leaf_of_node[u], e, fwd, n, valid, Ac);
        }
        // counters += (state not in A_c)
#pragma unroll
        for (int w = 0; w < W; w++) {
            uint32_t carry = ~Ac[w] & valid[w];
#pragma unroll
            for (int p = 0; p < P; p++) {
                const uint32_t s = plane[p][w] ^ carry;
                carry &= plane[p][w];
                plane[p][w] = s;
            }
        }
    }
    // states at the lowest count, and at the next one
    const int nch = c1 - c0;
    uint32_t Ab[W], Bb[W];
#pragma unroll
    for (int w = 0; w < W; w++) { Ab[w] = 0; Bb[w] = 0; }
    for (int k = 0; k <= nch; k++) {
        uint32_t any = 0, m[W];
#pragma unroll
        for (int w = 0; w < W; w++) {
            m[w] = valid[w];
#pragma unroll
            for (int p = 0; p < P; p++) m[w] &= ((k >> p) & 1) ? plane[p][w] : ~plane[p][w];
            any |= m[w];
        }
        if (any) {
#pragma unroll
            for (int w = 0; w < W; w++) {
                Ab[w] = m[w];
                uint32_t m1 = valid[w];
#pragma unroll
                for (int p = 0; p < P; p++) m1 &= (((k + 1) >> p) & 1) ? plane[p][w] : ~plane[p][w];
                Bb[w] = k + 1 <= nch ? m1 : 0u;
            }
            break;
        }
    }
#pragma unroll
    for (int w = 0; w < W; w++) {
        a.A[((size_t)ent * W + w) * ni + t] = Ab[w];
        a.B[((size_t)ent * W + w) * ni + t] = Bb[w];
    }
}

// backward step: all optimal states of node v given its parent's (already final, stored over the parent's B)
template <int W>
__device__ __forceinline__ void bep_node_backward(const BepArgs &a, int t, int v, int ent)
{
    const size_t ni = (size_t)a.n_inst;
    uint32_t Av[W], Bv[W], opt[W];
#pragma unroll
    for (int w = 0; w < W; w++) { Av[w] = a.A[((size_t)ent * W + w) * ni + t]; Bv[w] = a.B[((size_t)ent * W + w) * ni + t]; }
    const int par = a.parent[v];
    if (par < 0) {
#pragma unroll
        for (int w = 0; w < W; w++) opt[w] = Av[w];
    } else {
        const int pe = a.ent_of_node[par];
        uint32_t outside = 0, ob[W];
#pragma unroll
        for (int w = 0; w < W; w++) { ob[w] = a.B[((size_t)pe * W + w) * ni + t]; outside |= ob[w] & ~Av[w]; }
#pragma unroll
        for (int w = 0; w < W; w++) opt[w] = (ob[w] & (Av[w] | Bv[w])) | (outside ? Av[w] : 0u);
    }
#pragma unroll
    for (int w = 0; w < W; w++) a.B[((size_t)ent * W + w) * ni + t] = opt[w];
}

// Schedule 1 -- one thread walks the whole tree of its instance (any tree shape, two launches in total):
// children before parents going up (parent[v] < v), root first going down.
template <int W, int P>
__global__ void bep_forward_kernel(const BepArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_inst) return;
    const int e = a.inst_e[t];
    const bool fwd = a.inst_fwd[t] != 0;
    const int n = a.nsym[t];
    uint32_t valid[W];
    bep_valid_mask<W>(a, n, valid);
    for (int v = a.n_nodes - 1; v >= 0; v--) {
        const int ent = a.ent_of_node[v];
        if (ent >= 0) bep_node_forward<W, P>(a, t, v, ent, e, fwd, n, valid);
    }
}
template <int W>
__global__ void bep_backward_kernel(const BepArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_inst) return;
    for (int v = 0; v < a.n_nodes; v++) {
        const int ent = a.ent_of_node[v];
        if (ent >= 0) bep_node_backward<W>(a, t, v, ent);
    }
}

// Schedule 2 -- one launch per height level, one thread per (instance, node of the level): the children of a
// node sit on lower levels, its parent on a higher one.  (instance, node) parallelism instead of ~10,000 threads.
template <int W, int P>
__global__ void bep_forward_level_kernel(const BepArgs a, const int32_t *nodes, int n_level)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_inst) return;
    const int e = a.inst_e[t];
    const bool fwd = a.inst_fwd[t] != 0;
    const int n = a.nsym[t];
    uint32_t valid[W];
    bep_valid_mask<W>(a, n, valid);
    for (int k = blockIdx.y; k < n_level; k += gridDim.y) {
        const int v = nodes[k];
        bep_node_forward<W, P>(a, t, v, a.ent_of_node[v], e, fwd, n, valid);
    }
}
template <int W>
__global__ void bep_backward_level_kernel(const BepArgs a, const int32_t *nodes, int n_level)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_inst) return;
    for (int k = blockIdx.y; k < n_level; k += gridDim.y) {
        const int v = nodes[k];
        bep_node_backward<W>(a, t, v, a.ent_of_node[v]);
    }
}

// ---------------------------------------------------------------------------------------------------
// host-callable launchers (bep_host.cpp)
// ---------------------------------------------------------------------------------------------------
cudaError_t bep_launch_next_prev(const uint8_t *obs, const int32_t *leaf_nodes, int n_leaves, int n_pos, int32_t *nxt, int32_t *prv,
                                 cudaStream_t st)
{
    if (n_leaves == 0) return cudaSuccess;
    bep_next_prev_kernel<<<(n_leaves + 127) / 128, 128, 0, st>>>(obs, leaf_nodes, n_leaves, n_pos, nxt, prv);
    return cudaGetLastError();
}

cudaError_t bep_launch_dict(const BepArgs &a, cudaStream_t st)
{
    if (a.n_inst == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(a.symmap, 0, (size_t)a.n_inst * (a.n_pos + 2), st);
    if (e != cudaSuccess) return e;
    if (a.n_leaves > 0)
        bep_mark_kernel<<<dim3((a.n_inst + 127) / 128, a.n_leaves < 4096 ? a.n_leaves : 4096), 128, 0, st>>>(a);
    bep_number_kernel<<<(a.n_inst + 3) / 4, 128, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t bep_launch_sweeps(const BepArgs &a, int W, cudaStream_t st)
{
    if (a.n_inst == 0) return cudaSuccess;
    const int grid = (a.n_inst + 63) / 64;
    // counters of 6 bits serve every tree with at most 63 children per node; wider multifurcations take 20 bits
    const bool wide = a.max_children > 63;
#define BEP_FWD(W_) (wide ? bep_forward_kernel<W_, 20><<<grid, 64, 0, st>>>(a) : bep_forward_kernel<W_, 6><<<grid, 64, 0, st>>>(a))
    switch (W) {
    case 1: BEP_FWD(1); bep_backward_kernel<1><<<grid, 64, 0, st>>>(a); break;
    case 2: BEP_FWD(2); bep_backward_kernel<2><<<grid, 64, 0, st>>>(a); break;
    case 4: BEP_FWD(4); bep_backward_kernel<4><<<grid, 64, 0, st>>>(a); break;
    case 8: BEP_FWD(8); bep_backward_kernel<8><<<grid, 64, 0, st>>>(a); break;
    default: return cudaErrorInvalidValue;
    }
#undef BEP_FWD
    return cudaGetLastError();
}


template <int W>
static cudaError_t bep_levels_impl(const BepArgs &a, const int32_t *d_nodes, const int32_t *level_off, int n_levels, cudaStream_t st)
{
    const unsigned gx = (a.n_inst + 127) / 128;
    for (int h = 0; h < n_levels; h++) {
        const int cnt = level_off[h + 1] - level_off[h];
        if (cnt > 0) {
            const dim3 grid(gx, cnt < 32768 ? cnt : 32768);
            if (a.max_children > 63) bep_forward_level_kernel<W, 20><<<grid, 128, 0, st>>>(a, d_nodes + level_off[h], cnt);
            else bep_forward_level_kernel<W, 6><<<grid, 128, 0, st>>>(a, d_nodes + level_off[h], cnt);
        }
    }
    for (int h = n_levels - 1; h >= 0; h--) {
        const int cnt = level_off[h + 1] - level_off[h];
        if (cnt > 0) bep_backward_level_kernel<W><<<dim3(gx, cnt < 32768 ? cnt : 32768), 128, 0, st>>>(a, d_nodes + level_off[h], cnt);
    }
    return cudaGetLastError();
}

// d_nodes: internal nodes grouped by height (device), level_off: [n_levels + 1] offsets (host), lowest level first
cudaError_t bep_launch_sweeps_levels(const BepArgs &a, int W, const int32_t *d_nodes, const int32_t *level_off, int n_levels,
                                     cudaStream_t st)
{
    if (a.n_inst == 0) return cudaSuccess;
    switch (W) {
    case 1: return bep_levels_impl<1>(a, d_nodes, level_off, n_levels, st);
    case 2: return bep_levels_impl<2>(a, d_nodes, level_off, n_levels, st);
    case 4: return bep_levels_impl<4>(a, d_nodes, level_off, n_levels, st);
    case 8: return bep_levels_impl<8>(a, d_nodes, level_off, n_levels, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace bnk
