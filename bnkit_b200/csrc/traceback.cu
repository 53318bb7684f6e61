// traceback.cu -- max-product traceback (CGTable.getMPE, src/bn/alg/CGTable.java:885-948;
// DenseFactor.getAssign, src/bn/factor/DenseFactor.java:256-276) and small utility kernels.
//
// The reference follows per-factor pointer chains recursively; on a tree that is
//   state[v] = ptr_v[ state[parent(v)] ]     (ptr rows written by the joint up-sweep)
// walked root-to-leaves over the static pre-order program.
#include "device_common.cuh"

#include <cstdlib>
#include <cstring>

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
            if (kcol) kind = kcol[(size_t)(unsigned)__shfl_sync(FULL, c_ent, u) * ncols_u] & 3;  // (the masked warp-tiled walker keeps child flags above the kind)
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

// Column-owned traceback for long walks (r2).  The warp-per-column kernel above costs ~300 ns per step on a
// 10,000-level caterpillar (3.0 ms): per step a warp loads 20 useful bytes, runs two dependent shuffles and lane 0
// stores one byte into a sector of its own.  Here a warp owns 32 consecutive sorted columns, lane = column.  The
// walker wrote ptr as [ent][column][K] bytes, so the pointer bytes of a step for the warp's columns are ONE
// contiguous run of 32 K bytes: they stream into a shared-memory ring with cp.async TC_NB - 1 batches of TC_B
// steps ahead (16-byte copies -- two instructions per step -- when the run is 16-byte aligned and a multiple of
// 16 bytes, else 4-byte copies: a run is always 4-byte aligned), and the dependent chain of a step is one
// shared-memory byte load: the parent's state comes from a register when the parent was the previous step
// (program order is pre-order, so that is the common case; a uniform branch), else from a per-lane stack in
// shared memory.  States leave as one 32-byte row segment per step.  (First cut, 4-byte copies only and every
// field re-derived per step: 110 instructions per step and warp, 2.9 ms -- no faster than the kernel it replaced;
// ncu: one warp per scheduler, `wait` stalls, i.e. the instruction count IS the time.)
constexpr int TC_B = 16, TC_NB = 4, TC_WARPS = 2, TC_MAXSLOTS = 192;

template <bool HAS_KIND, bool HAS_OBS>
__global__ void __launch_bounds__(TC_WARPS * 32) traceback_cols_kernel(const TracebackArgs a)
{
    extern __shared__ __align__(16) uint8_t tc_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int K = a.K;
    const int col0 = a.col_lo + (blockIdx.x * TC_WARPS + wid) * 32;
    const int ncw = min(32, a.col_hi - col0);
    if (ncw <= 0) return;  // no CTA-wide barrier below
    const int slots = a.tb_slots > 0 ? a.tb_slots : 1;
    const int row_bytes = 32 * K;                         // one step of the ring
    const int warp_bytes = TC_NB * TC_B * row_bytes + slots * 32;
    uint8_t *ring = tc_smem + (size_t)wid * warp_bytes;
    uint8_t *stack = ring + TC_NB * TC_B * row_bytes;     // [slot][lane]
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    const int lc = lane < ncw ? lane : ncw - 1;           // spare lanes shadow the last column
    const int col = col0 + lc;
    const unsigned ncols_u = (unsigned)a.ncols;
    const int run = ncw * K;                              // bytes of one step's run
    const size_t ent_stride = (size_t)ncols_u * K;
    const uint8_t *pbase = a.ptr + (size_t)col0 * K;
    // 16-byte copies: every run starts on a 16-byte boundary and is a whole number of 16-byte pieces
    const bool wide = (run & 15) == 0 && (ent_stride & 15) == 0 && (reinterpret_cast<uintptr_t>(pbase) & 15) == 0;
    const int n_steps = a.n_steps, n_batches = (n_steps + TC_B - 1) / TC_B;
    const unsigned FULL = 0xffffffffu;
    const int ocol = (HAS_KIND && HAS_OBS) ? a.orig_col[col] : 0;
    uint8_t *srow = a.state_s + col;

    // copies of batch j: lane u < TC_B knows the step's ent; every lane copies its share of each step's run
    auto load_ent = [&](int j) -> int {
        const int si = j * TC_B + lane;
        return (lane < TC_B && si < n_steps) ? a.down[si].ent : -1;
    };
    auto issue = [&](int j, int my_ent) {
        if (j < n_batches) {
            const unsigned dst0 = ring_s + (unsigned)((j % TC_NB) * TC_B) * row_bytes;
            if (wide) {
                const int n16 = run >> 4;  // <= 2 K <= 40 pieces: at most two per lane
#pragma unroll
                for (int u = 0; u < TC_B; u++) {
                    const int ent = __shfl_sync(FULL, my_ent, u);  // < 0 past the end of the program: nothing to copy
                    const uint8_t *src = pbase + (size_t)(unsigned)ent * ent_stride + lane * 16;
                    const unsigned dst = dst0 + u * row_bytes + lane * 16;
                    if (ent >= 0 && lane < n16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
                    if (ent >= 0 && lane + 32 < n16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + 512), "l"(src + 512) : "memory");
                }
            } else {
                const int n_words = run >> 2;
#pragma unroll 2
                for (int u = 0; u < TC_B; u++) {
                    const int ent = __shfl_sync(FULL, my_ent, u);
                    if (ent < 0) break;  // uniform
                    const uint8_t *src = pbase + (size_t)(unsigned)ent * ent_stride;
                    for (int w = lane; w < n_words; w += 32)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst0 + u * row_bytes + w * 4), "l"(src + w * 4) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    // program fields of batch j (lane u carries step j * TC_B + u) and, HAS_KIND, this lane's kind bytes of the batch
    int f_node = 0, f_meta = 0;
    uint8_t f_kind[HAS_KIND ? TC_B : 1];
    auto fetch = [&](int j) {
        const int si = j * TC_B + lane;
        int my_ent = -1;
        if (lane < TC_B && si < n_steps) {
            const Step &sp = a.down[si];
            f_node = sp.node; my_ent = sp.ent;
            // 0..11 pslot + 1, 12..23 oslot + 1, 24: root step
            f_meta = (sp.pslot + 1) | ((sp.oslot + 1) << 12) | ((sp.parent < 0 ? 1 : 0) << 24);
        }
        if (HAS_KIND) {
#pragma unroll
            for (int u = 0; u < TC_B; u++) {
                const int ent = __shfl_sync(FULL, my_ent, u);
                f_kind[u] = ent >= 0 ? a.kind[(size_t)(unsigned)ent * ncols_u + col] : (uint8_t)KIND_NONE;
            }
        }
    };
    for (int j = 0; j < TC_NB - 1; j++) issue(j, load_ent(j));
    fetch(0);
    int next_ent = load_ent(TC_NB - 1);  // requested one batch before the copies that need it are issued
    int last_oslot = -2, last_s = 0;
    for (int b = 0; b < n_batches; b++) {
        const int c_node = f_node, c_meta = f_meta;
        uint8_t c_kind[HAS_KIND ? TC_B : 1];
        if (HAS_KIND) {
#pragma unroll
            for (int u = 0; u < TC_B; u++) c_kind[u] = f_kind[u];
        }
        __syncwarp();             // every lane has finished reading the ring slot batch b + TC_NB - 1 overwrites
        issue(b + TC_NB - 1, next_ent);
        next_ent = load_ent(b + TC_NB);
        if (b + 1 < n_batches) fetch(b + 1);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(TC_NB - 1) : "memory");
        __syncwarp();             // ... and every lane's copies of batch b have landed
        const uint8_t *rows = ring + (size_t)((b % TC_NB) * TC_B) * row_bytes + lc * K;
        const int left = n_steps - b * TC_B;
#pragma unroll
        for (int u = 0; u < TC_B; u++) {
            const bool act = u < left;  // (no break: the shuffles stay convergent; steps past the end store nothing)
            const int node = __shfl_sync(FULL, c_node, u), meta = __shfl_sync(FULL, c_meta, u);
            const int pslot = (meta & 0xfff) - 1, oslot = ((meta >> 12) & 0xfff) - 1;
            // the parent's state: the previous step's when the parent was the previous step (program order is
            // pre-order), else from the stack.  The stack read is unconditional and off the dependent chain -- its
            // address is known from the program; a stale value is never selected
            const int stv = stack[(pslot < 0 ? 0 : pslot) * 32 + lane];
            const int ps = pslot < 0 ? 0 : (pslot == last_oslot ? last_s : stv);
            int s;
            if (HAS_KIND) {
                const int kind = c_kind[u] & 3;  // (the masked warp-tiled walker keeps child flags above the kind)
                const int src = (kind == KIND_MSG && ps < K) ? ps : 0;
                s = rows[u * row_bytes + src];
                if (HAS_OBS && kind == KIND_OBS) s = a.obs[(size_t)(unsigned)node * ncols_u + ocol];
                else s = (kind == KIND_MSG || kind == KIND_ROOT) ? s : BNK_NONE;
            } else {
                s = rows[u * row_bytes + ((meta >> 24) ? 0 : ps)];  // a root step reads entry 0
            }
            const bool keep = act && oslot >= 0;
            if (keep) stack[oslot * 32 + lane] = (uint8_t)s;
            last_s = keep ? s : last_s;
            last_oslot = keep ? oslot : last_oslot;
            if (act && lane < ncw) srow[(size_t)(unsigned)node * ncols_u] = (uint8_t)s;
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

cudaError_t launch_traceback(const TracebackArgs &a, cudaStream_t st)
{
    if (a.tb_slots >= 4095) return cudaErrorInvalidValue;
    const char *tb_env = std::getenv("BNK_TRACEBACK");  // "warp": the warp-per-column kernel (its only other use: > TC_MAXSLOTS stack slots)
    const bool old_tb = tb_env && std::strcmp(tb_env, "warp") == 0;
    if (!old_tb && a.tb_slots <= TC_MAXSLOTS && a.n_steps > 0 && a.col_hi > a.col_lo && (a.K & 3) == 0) {
        const int n_warps = (a.col_hi - a.col_lo + 31) / 32;
        const unsigned grid = (n_warps + TC_WARPS - 1) / TC_WARPS;
        const size_t smem = (size_t)TC_WARPS * (TC_NB * TC_B * 32 * a.K + (a.tb_slots > 0 ? a.tb_slots : 1) * 32);
        auto go = [&](auto kern) -> cudaError_t {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return e;
            kern<<<grid, TC_WARPS * 32, smem, st>>>(a);
            return cudaGetLastError();
        };
        if (a.kind && a.kind_obs) return go(traceback_cols_kernel<true, true>);
        if (a.kind) return go(traceback_cols_kernel<true, false>);
        return go(traceback_cols_kernel<false, false>);
    }
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
#pragma unroll 4
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
    // a thread keeps its column and walks a slice of the leaves (8,192 slices were 160,000 CTAs of one or two bytes per
    // thread at C3: 0.18 ms for 50 MB)
    dim3 grid((a.ncols + 255) / 256, a.n_leaves < 256 ? a.n_leaves : 256);
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
