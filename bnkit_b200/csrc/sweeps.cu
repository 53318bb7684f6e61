// sweeps.cu -- the tree sweeps: max-product up-sweep (joint), sum-product up-sweep
// (inside / log-likelihood) and down-sweep (outside + posteriors), for sm_100a.
//
// What the reference does per column with generic bucket elimination
// (bn.alg.VarElim.infer, src/bn/alg/VarElim.java:187-334; Factorize.getProduct / getMaxMargin /
// getMargin, src/bn/factor/Factorize.java:632-1000, :1414-1543, :1223-1401) is, on a tree,
// leaf-to-root message passing.  Here one CTA owns a tile of alignment columns and walks the
// whole tree for them (program built in tree_plan.cpp): thread (c, p) owns entry p of the
// 20-state vectors of column c.  Live messages sit in a shared-memory stack; the node's
// 20x20 table (log P for max-product, P for sum-product, P^T for the down-sweep) is streamed
// into shared memory with cp.async three nodes deep; children that are alignment leaves
// never materialise a message -- their contribution is one table lookup by the observed
// state.  No launch per tree level, no message round trip through HBM for the joint pass.
//
// Arithmetic contracts
//   joint:  log space, adds only, in the reference's association order
//           non-root  L_v[p][x] + ((g1[x] + g2[x]) + g3[x] ...)      (Factorize.java:351-373)
//           root-like ((base[x] + g1[x]) + g2[x]) ...                base = log prior | L_v[obs_parent][x]
//           argmax = first strict maximum scanning x ascending, initial (-inf, 0)
//           (Factorize.java:1472-1481)  => states are bit-identical to the oracle's.
//   sum:    linear space FP64 with exact power-of-two rescaling per (node, column); the
//           exponents are summed per column and folded back into the log-likelihood.
//           Posteriors / log-likelihoods agree with the reference's log-space values to
//           rounding (tolerance 1e-9 relative in the tests).
#include "device_common.cuh"

namespace bnk {

constexpr int NSTAGE = 3;  // table ring depth (prefetch distance 2 nodes)
constexpr int MAXDEG_LOCAL = 16;

__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gsrc)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int K>
struct Smem {
    static constexpr int LROW = K + 2;  // padded row: conflict-free 128-bit row reads, 16 B aligned
    double *tab;    // [NSTAGE][ncat_max][K][LROW]
    double *gs;     // [2][tcpad][LROW]
    double *unit;   // [LROW]  zeros (log space) or ones (linear space)
    double *stack;  // [smem_slots][tcpad][K]
    __device__ Smem(unsigned char *raw, int ncat_max, int tcpad)
    {
        tab = reinterpret_cast<double *>(raw);
        gs = tab + (size_t)NSTAGE * ncat_max * K * LROW;
        unit = gs + (size_t)2 * tcpad * LROW;
        stack = unit + LROW;
    }
};

size_t walk_smem_bytes(int K, int cpt, int ncg, int ncat_max, int slots)
{
    const size_t lrow = K + 2, tcpad = (size_t)ncg * cpt;
    return sizeof(double) * ((size_t)NSTAGE * ncat_max * K * lrow + 2 * tcpad * lrow + lrow + (size_t)slots * tcpad * K);
}

int walk_max_smem_optin(int device)
{
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    return v;
}

// stream one node's row table (all categories of this tile) into a ring stage
template <int K>
__device__ __forceinline__ void prefetch_table(const WalkArgs &a, const TileDesc &tile, int node, double *stage_base)
{
    constexpr int LROW = K + 2, CH = K / 2;  // 16-byte chunks per row
    const int total = tile.ncat * K * CH;
    for (int q = threadIdx.x; q < total; q += blockDim.x) {
        const int ct = q / (K * CH), r = q - ct * (K * CH);
        const int row = r / CH, ch = r - row * CH;
        const double *src = a.T_row + ((size_t)(tile.cat0 + ct) * a.n_nodes + node) * (K * K) + row * K + ch * 2;
        cp_async_16(stage_base + ((size_t)ct * K + row) * LROW + ch * 2, src);
    }
}

__device__ __forceinline__ double pow2_scale_from_hi(int mhi, int &e)
{
    // mhi = largest high word among non-negative doubles; returns 2^-e with e = unbiased exponent of that maximum
    const int be = mhi >> 20;
    if (be == 0) { e = 0; return 1.0; }
    e = be - 1023;
    return __hiloint2double((2046 - be) << 20, 0);
}

// ------------------------------------------------------------------------------------
// up-sweeps.  MODE 0: max-product in log space (joint).  MODE 1: sum-product, linear space.
// ------------------------------------------------------------------------------------
template <int K, int CPT, bool GENERAL, int MODE>
__global__ void __launch_bounds__(768) up_kernel(const WalkArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LROW = K + 2;
    const TileDesc tile = a.tiles[blockIdx.x];
    const int tcpad = a.ncg * CPT;
    Smem<K> sm(smem_raw, a.ncat_max, tcpad);
    const int tid = threadIdx.x, p = tid % K, cg = tid / K;
    const size_t stage_elems = (size_t)a.ncat_max * K * LROW;

    int lc[CPT], col[CPT], tcat[CPT], ocol[CPT];
    bool act[CPT];
#pragma unroll
    for (int j = 0; j < CPT; j++) {
        lc[j] = cg * CPT + j;
        act[j] = (cg < a.ncg) && (lc[j] < tile.ncols);
        col[j] = tile.col0 + (act[j] ? lc[j] : 0);
        tcat[j] = a.col_cat[col[j]] - tile.cat0;
        ocol[j] = a.orig_col[col[j]];
    }
    for (int q = tid; q < LROW; q += blockDim.x) sm.unit[q] = (MODE == 0) ? 0.0 : 1.0;

    // per-column accumulators of the sum-product (identical in the K threads of a column)
    int esum[CPT];
    double logacc[CPT];
#pragma unroll
    for (int j = 0; j < CPT; j++) { esum[j] = 0; logacc[j] = 0.0; }

    // prologue: tables of the first two steps
    for (int s = 0; s < NSTAGE - 1; s++) {
        if (s < a.n_steps) prefetch_table<K>(a, tile, a.up[s].node, sm.tab + s * stage_elems);
        cp_async_commit();
    }

    for (int i = 0; i < a.n_steps; i++) {
        const UpStep st = a.up[i];
        const int v = st.node;
        double *gs = sm.gs + (size_t)(i & 1) * tcpad * LROW;
        uint8_t kind[CPT];
        // ---------------- fold the children ----------------
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            kind[j] = KIND_NONE;
            if (!act[j]) continue;
            const size_t tb = (size_t)(tile.cat0 + tcat[j]) * a.n_nodes;  // table row base of this column's category
            bool pv = true, pp = st.parent >= 0;
            uint8_t ov = BNK_NONE, op = BNK_NONE;
            if (GENERAL) {
                if (a.present) {
                    pv = a.present[(size_t)v * a.ncols + col[j]] != 0;
                    pp = pp && a.present[(size_t)st.parent * a.ncols + col[j]] != 0;
                }
                ov = a.obs[(size_t)v * a.ncols + col[j]];
                if (pp) op = a.obs[(size_t)st.parent * a.ncols + col[j]];
            }
            const bool rootlike = !pp || op != BNK_NONE;
            double G;
            if (MODE == 0) G = 0.0; else G = 1.0;
            bool first = true;
            if (rootlike) {
                G = (op != BNK_NONE) ? a.T_col[(tb + v) * (K * K) + p * K + op] : a.prior[p];
                first = false;
            }
            int nk = 0;
            const bool need = pv && ov == BNK_NONE;  // only unobserved present nodes combine messages
            for (int e = 0; e < st.n_child; e++) {
                const ChildRef ch = a.up_child[st.child_begin + e];
                if (GENERAL && a.present && a.present[(size_t)ch.node * a.ncols + col[j]] == 0) continue;
                nk++;
                if (!need) {
                    if (MODE == 1 && GENERAL && pv && ch.slot < 0) {  // observed v: scalar factors of its observed childless children
                        const uint8_t ou = a.obs[(size_t)ch.node * a.ncols + col[j]];
                        if (ou != BNK_NONE) logacc[j] += log(a.T_col[(tb + ch.node) * (K * K) + ou * K + ov]);
                    }
                    continue;
                }
                double g;
                uint8_t ou = BNK_NONE;
                if (ch.slot < 0 || GENERAL) ou = a.obs[(size_t)ch.node * a.ncols + col[j]];
                if (ou != BNK_NONE) {
                    g = a.T_col[(tb + ch.node) * (K * K) + ou * K + p];
                } else if (ch.slot < 0) {  // unobserved childless node: marginalise / maximise its own row
                    const double *row = a.T_row + (tb + ch.node) * (K * K) + p * K;
                    g = row[0];
                    for (int x = 1; x < K; x++) g = (MODE == 0) ? (row[x] > g ? row[x] : g) : g + row[x];
                } else if (ch.slot < a.smem_slots) {
                    g = sm.stack[((size_t)ch.slot * tcpad + lc[j]) * K + p];
                } else {
                    g = a.spill[((size_t)(ch.slot - a.smem_slots) * a.ncols + col[j]) * K + p];
                }
                if (MODE == 0) G = first ? g : G + g;
                else G = first ? g : G * g;
                first = false;
            }
            if (!pv) kind[j] = KIND_NONE;
            else if (ov != BNK_NONE) kind[j] = KIND_OBS;
            else if (!pp && nk == 0) kind[j] = KIND_NONE;  // not connected: no BN node (PhyloBN.java:149)
            else kind[j] = rootlike ? KIND_ROOT : KIND_MSG;
            if (MODE == 1 && GENERAL && kind[j] == KIND_OBS) {  // scalar factor of an observed node
                if (!pp) { if (nk > 0) logacc[j] += log(a.prior[ov]); }
                else if (op != BNK_NONE) logacc[j] += log(a.T_col[(tb + v) * (K * K) + ov * K + op]);
            }
            gs[(size_t)lc[j] * LROW + p] = G;
        }
        cp_async_wait<NSTAGE - 2>();
        __syncthreads();
        if (i + NSTAGE - 1 < a.n_steps)
            prefetch_table<K>(a, tile, a.up[i + NSTAGE - 1].node, sm.tab + ((i + NSTAGE - 1) % NSTAGE) * stage_elems);
        cp_async_commit();
        // ---------------- contract with the node's table ----------------
        const double *stage = sm.tab + (i % NSTAGE) * stage_elems;
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            if (!act[j]) continue;
            if (MODE == 0 && p == 0) a.kind[(size_t)st.ent * a.ncols + ocol[j]] = kind[j];
            if (MODE == 1 && GENERAL && p == 0 && a.kindm) a.kindm[(size_t)st.ent * a.ncols + col[j]] = kind[j];
            if (kind[j] != KIND_MSG && kind[j] != KIND_ROOT) continue;
            const double2 *row2 = reinterpret_cast<const double2 *>(
                kind[j] == KIND_ROOT ? sm.unit : stage + ((size_t)tcat[j] * K + p) * LROW);
            const double2 *g2 = reinterpret_cast<const double2 *>(gs + (size_t)lc[j] * LROW);
            if (MODE == 0) {
                double best = -CUDART_INF;
                int bi = 0;
#pragma unroll
                for (int x2 = 0; x2 < K / 2; x2++) {
                    const double2 l = row2[x2], g = g2[x2];
                    const double v0 = l.x + g.x, v1 = l.y + g.y;
                    if (v0 > best) { best = v0; bi = 2 * x2; }
                    if (v1 > best) { best = v1; bi = 2 * x2 + 1; }
                }
                a.ptr[((size_t)st.ent * a.ncols + ocol[j]) * K + p] = (uint8_t)bi;
                if (st.slot >= 0) {
                    if (st.slot < a.smem_slots) sm.stack[((size_t)st.slot * tcpad + lc[j]) * K + p] = best;
                    else a.spill[((size_t)(st.slot - a.smem_slots) * a.ncols + col[j]) * K + p] = best;
                }
            } else {
                double acc0 = 0.0, acc1 = 0.0;
                int mhi = 0;
#pragma unroll
                for (int x2 = 0; x2 < K / 2; x2++) {
                    const double2 l = row2[x2], g = g2[x2];
                    acc0 = fma(l.x, g.x, acc0);
                    acc1 = fma(l.y, g.y, acc1);
                    mhi = max(mhi, max(__double2hiint(g.x), __double2hiint(g.y)));
                }
                int e;
                const double scale = pow2_scale_from_hi(mhi, e);
                const double m = (acc0 + acc1) * scale;
                esum[j] += e;
                if (kind[j] == KIND_ROOT) {
                    logacc[j] += log(m);
                } else {
                    if (st.slot >= 0) {
                        if (st.slot < a.smem_slots) sm.stack[((size_t)st.slot * tcpad + lc[j]) * K + p] = m;
                        else a.spill[((size_t)(st.slot - a.smem_slots) * a.ncols + col[j]) * K + p] = m;
                    }
                    if (a.msg) a.msg[((size_t)st.ent * a.ncols + col[j]) * K + p] = m;
                }
            }
        }
    }
    if (MODE == 1 && a.out_logl && p == 0) {
#pragma unroll
        for (int j = 0; j < CPT; j++)
            if (act[j]) a.out_logl[ocol[j]] = logacc[j] + (double)esum[j] * 0.693147180559945309417232121458;
    }
}

// ------------------------------------------------------------------------------------
// down-sweep: from-above vectors and posteriors of every internal node, pre-order.
//   D_v[x]   = sum_xp P_v[xp][x] * T_v[xp]          (T_v left on the stack by the parent's step)
//   post_v   = normalise(D_v * prod_children g_c)    (CGTable.query + EnumDistrib.normalise)
//   T_c[x]   = D_v[x] * prod_{w != c} g_w[x]         pushed for every internal child c
// ------------------------------------------------------------------------------------
template <int K, int CPT, bool GENERAL>
__global__ void __launch_bounds__(768) down_kernel(const WalkArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LROW = K + 2;
    const TileDesc tile = a.tiles[blockIdx.x];
    const int tcpad = a.ncg * CPT;
    Smem<K> sm(smem_raw, a.ncat_max, tcpad);
    const int tid = threadIdx.x, p = tid % K, cg = tid / K;
    const size_t stage_elems = (size_t)a.ncat_max * K * LROW;

    int lc[CPT], col[CPT], tcat[CPT], ocol[CPT];
    bool act[CPT];
#pragma unroll
    for (int j = 0; j < CPT; j++) {
        lc[j] = cg * CPT + j;
        act[j] = (cg < a.ncg) && (lc[j] < tile.ncols);
        col[j] = tile.col0 + (act[j] ? lc[j] : 0);
        tcat[j] = a.col_cat[col[j]] - tile.cat0;
        ocol[j] = a.orig_col[col[j]];
    }
    // ring: step i uses stage i % NSTAGE; the table of step i + NSTAGE is requested right after
    // the barrier of step i, when nobody reads stage i % NSTAGE any more.
    for (int s = 0; s < NSTAGE; s++) {
        if (s < a.n_steps) prefetch_table<K>(a, tile, a.down[s].node, sm.tab + s * stage_elems);
        cp_async_commit();
    }
    cp_async_wait<NSTAGE - 1>();
    __syncthreads();
    for (int i = 0; i < a.n_steps; i++) {
        const DownStep st = a.down[i];
        const int v = st.node;
        double *gs = sm.gs + (size_t)(i & 1) * tcpad * LROW;
        const double *stage = sm.tab + (i % NSTAGE) * stage_elems;
        uint8_t kind[CPT];
        double D[CPT], U[CPT];
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            kind[j] = KIND_NONE; D[j] = 0.0; U[j] = 0.0;
            if (!act[j]) continue;
            const size_t tb = (size_t)(tile.cat0 + tcat[j]) * a.n_nodes;
            kind[j] = GENERAL ? a.kindm[(size_t)st.ent * a.ncols + col[j]] : (st.parent < 0 ? KIND_ROOT : KIND_MSG);
            if (kind[j] != KIND_MSG && kind[j] != KIND_ROOT) continue;
            // ---- from-above vector ----
            if (kind[j] == KIND_ROOT) {
                uint8_t op = BNK_NONE;
                if (GENERAL && st.parent >= 0 && (!a.present || a.present[(size_t)st.parent * a.ncols + col[j]]))
                    op = a.obs[(size_t)st.parent * a.ncols + col[j]];
                D[j] = (op != BNK_NONE) ? a.T_col[(tb + v) * (K * K) + p * K + op] : a.prior[p];
            } else {
                const double2 *row2 = reinterpret_cast<const double2 *>(stage + ((size_t)tcat[j] * K + p) * LROW);
                double acc0 = 0.0, acc1 = 0.0;
                int mhi = 0;
                if (st.tslot < a.smem_slots) {
                    const double2 *t2 = reinterpret_cast<const double2 *>(sm.stack + ((size_t)st.tslot * tcpad + lc[j]) * K);
#pragma unroll
                    for (int x2 = 0; x2 < K / 2; x2++) {
                        const double2 l = row2[x2], t = t2[x2];
                        acc0 = fma(l.x, t.x, acc0);
                        acc1 = fma(l.y, t.y, acc1);
                        mhi = max(mhi, max(__double2hiint(t.x), __double2hiint(t.y)));
                    }
                } else {
                    const double2 *t2 = reinterpret_cast<const double2 *>(
                        a.spill + ((size_t)(st.tslot - a.smem_slots) * a.ncols + col[j]) * K);
#pragma unroll
                    for (int x2 = 0; x2 < K / 2; x2++) {
                        const double2 l = row2[x2], t = t2[x2];
                        acc0 = fma(l.x, t.x, acc0);
                        acc1 = fma(l.y, t.y, acc1);
                        mhi = max(mhi, max(__double2hiint(t.x), __double2hiint(t.y)));
                    }
                }
                int e;
                const double scale = pow2_scale_from_hi(mhi, e);
                D[j] = (acc0 + acc1) * scale;
            }
        }
        // child slots never alias this step's own slot (tree_plan.cpp allocates them before
        // releasing it), so they can be written without waiting for the other threads' reads
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            if (!act[j] || (kind[j] != KIND_MSG && kind[j] != KIND_ROOT)) continue;
            const size_t tb = (size_t)(tile.cat0 + tcat[j]) * a.n_nodes;
            // ---- children: messages, product, and the from-above vectors they receive ----
            double gv[MAXDEG_LOCAL];
            uint8_t push[MAXDEG_LOCAL];
            double G = 1.0;
            const int nch = st.n_child < MAXDEG_LOCAL ? st.n_child : MAXDEG_LOCAL;
            for (int e = 0; e < nch; e++) {
                const ChildRef ch = a.down_child[st.child_begin + e];
                gv[e] = 1.0; push[e] = 0;
                if (GENERAL && a.present && a.present[(size_t)ch.node * a.ncols + col[j]] == 0) continue;
                uint8_t ou = BNK_NONE;
                if (ch.ent < 0 || GENERAL) ou = a.obs[(size_t)ch.node * a.ncols + col[j]];
                if (ou != BNK_NONE) {
                    gv[e] = a.T_col[(tb + ch.node) * (K * K) + ou * K + p];
                } else if (ch.ent < 0) {
                    double s = 0.0;  // unobserved childless node: sum_x' P_u[p][x'] = sum_x' PT_u[x'][p]
                    for (int x = 0; x < K; x++) s += a.T_col[(tb + ch.node) * (K * K) + x * K + p];
                    gv[e] = s;
                } else {
                    gv[e] = a.msg[((size_t)ch.ent * a.ncols + col[j]) * K + p];
                    push[e] = 1;
                }
                G *= gv[e];
            }
            U[j] = D[j] * G;
            gs[(size_t)lc[j] * LROW + p] = U[j];
            // prefix / suffix products give prod_{w != c} without dividing
            double pre = D[j];
            double suf[MAXDEG_LOCAL];
            {
                double s = 1.0;
                for (int e = nch - 1; e >= 0; e--) { suf[e] = s; s *= gv[e]; }
            }
            for (int e = 0; e < nch; e++) {
                if (push[e]) {
                    const ChildRef ch = a.down_child[st.child_begin + e];
                    const double t = pre * suf[e];
                    if (ch.slot < a.smem_slots) sm.stack[((size_t)ch.slot * tcpad + lc[j]) * K + p] = t;
                    else a.spill[((size_t)(ch.slot - a.smem_slots) * a.ncols + col[j]) * K + p] = t;
                }
                pre *= gv[e];
            }
        }
        cp_async_wait<NSTAGE - 2>();  // table of step i + 1 has landed (this thread's copies)
        __syncthreads();              // ... and everybody's; gs and the child slots are visible
        if (i + NSTAGE < a.n_steps)
            prefetch_table<K>(a, tile, a.down[i + NSTAGE].node, sm.tab + (i % NSTAGE) * stage_elems);
        cp_async_commit();
        // ---- normalise and emit the posterior ----
        const int q = a.qrow ? a.qrow[st.ent] : st.ent;
        if (q >= 0 && a.out_post) {
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                if (!act[j]) continue;
                double out = CUDART_NAN;
                if (kind[j] == KIND_MSG || kind[j] == KIND_ROOT) {
                    const double2 *g2 = reinterpret_cast<const double2 *>(gs + (size_t)lc[j] * LROW);
                    double s = 0.0;
#pragma unroll
                    for (int x2 = 0; x2 < K / 2; x2++) { const double2 g = g2[x2]; s += g.x; s += g.y; }
                    out = U[j] / s;
                }
                a.out_post[((size_t)q * a.ncols + ocol[j]) * K + p] = out;
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------
template <typename KernelT>
static cudaError_t launch_walk(KernelT kern, const WalkLaunch &wl, const WalkArgs &a)
{
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wl.smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<wl.n_tiles, wl.threads, wl.smem_bytes, wl.stream>>>(a);
    return cudaGetLastError();
}

#define DISPATCH_WALK(KERNEL, ...)                                                              \
    do {                                                                                        \
        if (wl.K == 20) {                                                                       \
            if (wl.cpt == 1) return wl.general ? launch_walk(KERNEL<20, 1, true __VA_ARGS__>, wl, a)  \
                                               : launch_walk(KERNEL<20, 1, false __VA_ARGS__>, wl, a); \
            if (wl.cpt == 2) return wl.general ? launch_walk(KERNEL<20, 2, true __VA_ARGS__>, wl, a)  \
                                               : launch_walk(KERNEL<20, 2, false __VA_ARGS__>, wl, a); \
            if (wl.cpt == 4) return wl.general ? launch_walk(KERNEL<20, 4, true __VA_ARGS__>, wl, a)  \
                                               : launch_walk(KERNEL<20, 4, false __VA_ARGS__>, wl, a); \
        } else if (wl.K == 4) {                                                                 \
            if (wl.cpt == 1) return wl.general ? launch_walk(KERNEL<4, 1, true __VA_ARGS__>, wl, a)   \
                                               : launch_walk(KERNEL<4, 1, false __VA_ARGS__>, wl, a);  \
        }                                                                                       \
        return cudaErrorInvalidValue;                                                           \
    } while (0)

#define COMMA_ARG(x) , x
cudaError_t launch_joint_up(const WalkLaunch &wl, const WalkArgs &a) { DISPATCH_WALK(up_kernel, COMMA_ARG(0)); }
cudaError_t launch_sum_up(const WalkLaunch &wl, const WalkArgs &a) { DISPATCH_WALK(up_kernel, COMMA_ARG(1)); }
cudaError_t launch_sum_down(const WalkLaunch &wl, const WalkArgs &a) { DISPATCH_WALK(down_kernel); }

}  // namespace bnk
