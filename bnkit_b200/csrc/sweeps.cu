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
// into shared memory with cp.async three nodes deep, the walk program itself in 32-step
// chunks; children that are alignment leaves never materialise a message -- their
// contribution is one table lookup by the observed state, software-pipelined two steps
// ahead (state bytes at distance 2, table entries at distance 1) so that no global-memory
// latency sits on the per-node critical path.  No launch per tree level, no message round
// trip through HBM for the joint pass.
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

constexpr int NSTAGE = 3;   // table ring depth
constexpr int PCHUNK = 32;  // walk-program steps per shared-memory chunk (double buffered)
constexpr int MAXDEG_LOCAL = 16;
constexpr int MAXQ = 32;  // joint: at most MAXQ - 1 children per node in the general kernels

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
    Step *prog;     // [2][PCHUNK]
    double *tab;    // [NSTAGE][ncat_max][K][LROW]
    double *gs;     // [2][tcpad][LROW]
    double *unit;   // [LROW]  zeros (log space) or ones (linear space)
    double *stack;  // [smem_slots][tcpad][K]
    __device__ Smem(unsigned char *raw, int ncat_max, int tcpad)
    {
        prog = reinterpret_cast<Step *>(raw);
        tab = reinterpret_cast<double *>(raw + 2 * PCHUNK * sizeof(Step));
        gs = tab + (size_t)NSTAGE * ncat_max * K * LROW;
        unit = gs + (size_t)2 * tcpad * LROW;
        stack = unit + LROW;
    }
};

size_t walk_smem_bytes(int K, int cpt, int ncg, int ncat_max, int slots)
{
    const size_t lrow = K + 2, tcpad = (size_t)ncg * cpt;
    return 2 * PCHUNK * sizeof(Step) +
           sizeof(double) * ((size_t)NSTAGE * ncat_max * K * lrow + 2 * tcpad * lrow + lrow + (size_t)slots * tcpad * K);
}

int walk_max_smem_optin(int device)
{
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    return v;
}

// stream one node's row table (all categories of this tile) into a ring stage
template <int K>
__device__ __forceinline__ void prefetch_table(const double *T_row, int n_nodes, const TileDesc &tile, int node,
                                               double *stage_base)
{
    constexpr int LROW = K + 2, CH = K / 2;  // 16-byte chunks per row
    const int total = tile.ncat * K * CH;
    for (int q = threadIdx.x; q < total; q += blockDim.x) {
        const int ct = q / (K * CH), r = q - ct * (K * CH);
        const int row = r / CH, ch = r - row * CH;
        const double *src = T_row + ((size_t)(tile.cat0 + ct) * n_nodes + node) * (K * K) + row * K + ch * 2;
        cp_async_16(stage_base + ((size_t)ct * K + row) * LROW + ch * 2, src);
    }
}

__device__ __forceinline__ void prefetch_prog(const Step *prog_g, int n_steps, int chunk, Step *dst)
{
    const int q = threadIdx.x;  // one 16-byte piece per thread; blockDim.x >= 32, so loop
    for (int k = q; k < PCHUNK * 4; k += blockDim.x) {
        const int s = chunk * PCHUNK + (k >> 2);
        if (s < n_steps)
            cp_async_16(reinterpret_cast<char *>(dst) + k * 16, reinterpret_cast<const char *>(prog_g + s) + (k & 3) * 16);
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

// state bytes one step needs for one column (loaded two steps ahead)
struct Bytes {
    uint8_t ov, op;    // observation at the node / at its parent (BNK_NONE = none)
    uint8_t pv, pp;    // node present / parent present
    uint8_t ou[2];     // observations at the first two children
    uint8_t pu[2];     // their presence
};
// table entries one step needs for one column (loaded one step ahead)
struct Looks {
    double base;       // root-like start value: prior[p] or T[v][obs_parent -> p]
    double g[2];       // messages of observed first-two children: T[child][p -> obs_child]
};

template <bool GENERAL>
__device__ __forceinline__ Bytes load_bytes(const WalkArgs &a, const Step &s, int col /* ORIGINAL column */, bool down)
{
    Bytes b;
    b.ov = BNK_NONE; b.op = BNK_NONE; b.pv = 1; b.pp = s.parent >= 0;
    b.ou[0] = b.ou[1] = BNK_NONE; b.pu[0] = b.pu[1] = 1;
    if (GENERAL) {
        if (a.present) {
            b.pv = a.present[(size_t)s.node * a.ncols + col];
            if (s.parent >= 0) b.pp = a.present[(size_t)s.parent * a.ncols + col];
        }
        b.ov = a.obs[(size_t)s.node * a.ncols + col];
        if (s.parent >= 0) b.op = a.obs[(size_t)s.parent * a.ncols + col];
    }
#pragma unroll
    for (int e = 0; e < 2; e++) {
        if (e < s.n_child) {
            const bool childless = down ? (s.c_ent[e] < 0) : (s.c_slot[e] < 0);
            if (GENERAL && a.present) b.pu[e] = a.present[(size_t)s.c_node[e] * a.ncols + col];
            if (GENERAL || childless) b.ou[e] = a.obs[(size_t)s.c_node[e] * a.ncols + col];
        }
    }
    if (GENERAL && !b.pp) b.op = BNK_NONE;
    return b;
}

// ------------------------------------------------------------------------------------
// up-sweeps.  MODE 0: max-product in log space (joint).  MODE 1: sum-product, linear space.
// ------------------------------------------------------------------------------------
template <int K, int CPT, bool GENERAL, int MODE>
__global__ void __launch_bounds__(CPT == 1 ? 768 : (CPT == 2 ? 512 : 256)) up_kernel(const WalkArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LROW = K + 2, KK = K * K;
    const TileDesc tile = a.tiles[blockIdx.x];
    const int tcpad = a.ncg * CPT;
    Smem<K> sm(smem_raw, a.ncat_max, tcpad);
    const int tid = threadIdx.x, p = tid % K, cg = tid / K;
    const size_t stage_elems = (size_t)a.ncat_max * K * LROW;
    const int n_steps = a.n_steps;
    const double ident = (MODE == 0) ? 0.0 : 1.0;

    int lc[CPT], col[CPT], tcat[CPT], ocol[CPT];
    size_t tb[CPT];
    bool act[CPT];
#pragma unroll
    for (int j = 0; j < CPT; j++) {
        lc[j] = cg * CPT + j;
        act[j] = (cg < a.ncg) && (lc[j] < tile.ncols);
        col[j] = tile.col0 + (act[j] ? lc[j] : 0);
        tcat[j] = a.col_cat[col[j]] - tile.cat0;
        ocol[j] = a.orig_col[col[j]];
        tb[j] = (size_t)(tile.cat0 + tcat[j]) * a.n_nodes;
    }
    const double prior_p = a.prior[p];
    for (int q = tid; q < LROW; q += blockDim.x) sm.unit[q] = ident;

    // per-column accumulators of the sum-product (identical in the K threads of a column)
    int esum[CPT];
    double logacc[CPT];
#pragma unroll
    for (int j = 0; j < CPT; j++) { esum[j] = 0; logacc[j] = 0.0; }

    // prologue: program chunk 0, then the tables of the first two steps
    prefetch_prog(a.up, n_steps, 0, sm.prog);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    for (int s = 0; s < NSTAGE - 1; s++) {
        if (s < n_steps) prefetch_table<K>(a.T_row, a.n_nodes, tile, sm.prog[s].node, sm.tab + s * stage_elems);
        cp_async_commit();
    }
    auto step_ref = [&](int i) -> const Step & { return sm.prog[((i / PCHUNK) & 1) * PCHUNK + (i % PCHUNK)]; };
    auto load_looks = [&](const Step &s, const Bytes &b, int j) -> Looks {
        Looks l;
        l.base = ident; l.g[0] = l.g[1] = ident;
        const bool rootlike = !b.pp || b.op != BNK_NONE;
        if (rootlike) l.base = (b.op != BNK_NONE) ? a.T_col[(tb[j] + s.node) * KK + p * K + b.op] : prior_p;
#pragma unroll
        for (int e = 0; e < 2; e++)
            if (e < s.n_child && b.pu[e] && b.ou[e] != BNK_NONE)
                l.g[e] = a.T_col[(tb[j] + s.c_node[e]) * KK + b.ou[e] * K + p];
        return l;
    };

    Bytes B0[CPT], B1[CPT];
    Looks L0[CPT];
#pragma unroll
    for (int j = 0; j < CPT; j++) {
        if (act[j] && n_steps > 0) B0[j] = load_bytes<GENERAL>(a, step_ref(0), ocol[j], false);
        if (act[j] && n_steps > 1) B1[j] = load_bytes<GENERAL>(a, step_ref(1), ocol[j], false);
        if (act[j] && n_steps > 0) L0[j] = load_looks(step_ref(0), B0[j], j);
    }

    for (int i = 0; i < n_steps; i++) {
        const Step st = step_ref(i);
        const int v = st.node;
        double *gs = sm.gs + (size_t)(i & 1) * tcpad * LROW;
        uint8_t kind[CPT];
        Bytes B2[CPT];
        Looks L1[CPT];
        // ---------------- software pipeline: bytes of step i+2, table entries of step i+1 ----------------
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            if (!act[j]) continue;
            if (i + 2 < n_steps) B2[j] = load_bytes<GENERAL>(a, step_ref(i + 2), ocol[j], false);
            if (i + 1 < n_steps) L1[j] = load_looks(step_ref(i + 1), B1[j], j);
        }
        // ---------------- fold the children ----------------
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            kind[j] = KIND_NONE;
            if (!act[j]) continue;
            const Bytes b = B0[j];
            const bool pv = b.pv != 0, pp = b.pp != 0;
            const uint8_t ov = b.ov, op = b.op;
            const bool rootlike = !pp || op != BNK_NONE;
            double G = rootlike ? L0[j].base : ident;
            bool first = !rootlike;
            int nk = 0;
            const bool need = pv && ov == BNK_NONE;  // only unobserved present nodes combine messages
            // value of one present child's message for entry p of this column; the first two children use
            // the values prefetched a step ago
            auto child_g = [&](int cnode, int cslot, int cent, uint8_t ou, bool pref, double gpref) -> double {
                if (ou != BNK_NONE) return pref ? gpref : a.T_col[(tb[j] + cnode) * KK + ou * K + p];
                if (cslot == SLOT_WIDE)  // handled by a level kernel: message in HBM, [ent][K][ncols]
                    return a.msg[((size_t)cent * K + p) * a.ncols + col[j]];
                if (cslot < 0) {  // unobserved childless node: maximise / marginalise its own row
                    const double *row = a.T_row + (tb[j] + cnode) * KK + p * K;
                    double g = row[0];
                    for (int x = 1; x < K; x++) g = (MODE == 0) ? (row[x] > g ? row[x] : g) : g + row[x];
                    return g;
                }
                if (cslot < a.smem_slots) return sm.stack[((size_t)cslot * tcpad + lc[j]) * K + p];
                return a.spill[((size_t)(cslot - a.smem_slots) * a.ncols + col[j]) * K + p];
            };
            auto child_info = [&](int e, int &cnode, int &cslot, int &cent, bool &pu, uint8_t &ou) {
                if (e < 2) {
                    cnode = st.c_node[e]; cslot = st.c_slot[e]; cent = st.c_ent[e];
                    pu = (e == 0 ? b.pu[0] : b.pu[1]) != 0; ou = (e == 0) ? b.ou[0] : b.ou[1];
                } else {
                    const ChildRef ch = a.up_child[st.child_begin + e];
                    cnode = ch.node; cslot = ch.slot; cent = ch.ent;
                    pu = !(GENERAL && a.present) || a.present[(size_t)cnode * a.ncols + ocol[j]] != 0;
                    ou = (GENERAL || cslot < 0) ? a.obs[(size_t)cnode * a.ncols + ocol[j]] : (uint8_t)BNK_NONE;
                }
            };
            if (MODE == 0 && st.n_child > 2 && need) {
                // >= 3 single-variable factors: the reference multiplies them along Factorize.getProductTree
                // (Factorize.java:331-375) = queue pairing over [base, observed children, unobserved children]
                double qloc[2 * MAXQ + 1];
                double *q = qloc;  // nodes with MAXQ or more children keep the queue in global scratch (any arity)
                if (st.n_child >= MAXQ) q = a.qscratch + ((size_t)(blockIdx.x * blockDim.x + tid) * CPT + j) * a.qstride;
                int m = 0;
                if (rootlike) q[m++] = G;
                for (int pass = 0; pass < 2; pass++) {
                    for (int e = 0; e < st.n_child; e++) {
                        int cnode, cslot, cent; bool pu; uint8_t ou;
                        child_info(e, cnode, cslot, cent, pu, ou);
                        if (!pu) continue;
                        if (pass == 0) nk++;
                        if ((ou != BNK_NONE) != (pass == 0)) continue;
                        q[m++] = child_g(cnode, cslot, cent, ou, e < 2, e == 0 ? L0[j].g[0] : L0[j].g[1]);
                    }
                }
                int head = 0, tail = m;
                while (tail - head > 1) {
                    const double x0 = q[head++], x1 = q[head++];
                    q[tail++] = x0 + x1;
                }
                G = m > 0 ? q[head] : 0.0;
            } else {
                // root-like nodes combine [base, observed children, unobserved children]; with two children
                // of mixed status that order differs from index order in the last bit, so take two passes
                const int passes = (MODE == 0 && rootlike && st.n_child == 2) ? 2 : 1;
                for (int pass = 0; pass < passes; pass++)
                for (int e = 0; e < st.n_child; e++) {
                    int cnode, cslot, cent; bool pu; uint8_t ou;
                    child_info(e, cnode, cslot, cent, pu, ou);
                    if (!pu) continue;
                    if (passes == 2 && (ou != BNK_NONE) != (pass == 0)) continue;
                    nk++;
                    if (!need) {
                        if (MODE == 1 && GENERAL && pv && cslot == SLOT_CHILDLESS && ou != BNK_NONE)  // observed v: scalar factor of an observed childless child
                            logacc[j] += log(a.T_col[(tb[j] + cnode) * KK + ou * K + ov]);
                        continue;
                    }
                    const double g = child_g(cnode, cslot, cent, ou, e < 2, e == 0 ? L0[j].g[0] : L0[j].g[1]);
                    if (MODE == 0) G = first ? g : G + g;
                    else G = first ? g : G * g;
                    first = false;
                }
            }
            if (!pv) kind[j] = KIND_NONE;
            else if (ov != BNK_NONE) kind[j] = KIND_OBS;
            else if (!pp && nk == 0) kind[j] = KIND_NONE;  // not connected: no BN node (PhyloBN.java:149)
            else kind[j] = rootlike ? KIND_ROOT : KIND_MSG;
            if (MODE == 1 && GENERAL && kind[j] == KIND_OBS) {  // scalar factor of an observed node
                if (!pp) { if (nk > 0) logacc[j] += log(a.prior[ov]); }
                else if (op != BNK_NONE) logacc[j] += log(a.T_col[(tb[j] + v) * KK + ov * K + op]);
            }
            gs[(size_t)lc[j] * LROW + p] = G;
        }
        cp_async_wait<NSTAGE - 2>();
        __syncthreads();
        if (i + NSTAGE - 1 < n_steps)
            prefetch_table<K>(a.T_row, a.n_nodes, tile, step_ref(i + NSTAGE - 1).node,
                              sm.tab + ((i + NSTAGE - 1) % NSTAGE) * stage_elems);
        if (i % PCHUNK == 0) prefetch_prog(a.up, n_steps, i / PCHUNK + 1, sm.prog + (((i / PCHUNK) + 1) & 1) * PCHUNK);
        cp_async_commit();
        // ---------------- contract with the node's table ----------------
        const double *stage = sm.tab + (i % NSTAGE) * stage_elems;
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            if (!act[j]) continue;
            if (MODE == 0 && p == 0) a.kind[(size_t)st.ent * a.ncols + col[j]] = kind[j];
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
                a.ptr[((size_t)st.ent * a.ncols + col[j]) * K + p] = (uint8_t)bi;
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
#pragma unroll
        for (int j = 0; j < CPT; j++) { B0[j] = B1[j]; B1[j] = B2[j]; L0[j] = L1[j]; }
    }
    if (MODE == 1 && a.out_logl && p == 0) {
#pragma unroll
        for (int j = 0; j < CPT; j++)
            if (act[j]) {
                const int ew = a.esum_wide ? a.esum_wide[col[j]] : 0;
                const double lw = a.logl_wide ? a.logl_wide[col[j]] : 0.0;
                a.out_logl[ocol[j]] = logacc[j] + lw + (double)(esum[j] + ew) * 0.693147180559945309417232121458;
            }
    }
}

// ------------------------------------------------------------------------------------
// down-sweep: from-above vectors and posteriors of every internal node, pre-order.
//   D_v[x]   = sum_xp P_v[xp][x] * T_v[xp]          (T_v left on the stack by the parent's step)
//   post_v   = normalise(D_v * prod_children g_c)    (CGTable.query + EnumDistrib.normalise)
//   T_c[x]   = D_v[x] * prod_{w != c} g_w[x]         pushed for every internal child c
// ------------------------------------------------------------------------------------
// value of one of the first two children in the down-sweep: prefetched (observed child: table
// entry; internal child: stored up-message) or, for an uninstantiated childless node, the row sum
__device__ __forceinline__ double child_value(const WalkArgs &a, const Step &st, const Bytes &b, const Looks &l, int e,
                                              size_t tb, int col, int p, int K)
{
    if (b.ou[e] != BNK_NONE || st.c_ent[e] >= 0) return l.g[e];
    double s = 0.0;
    for (int x = 0; x < K; x++) s += a.T_col[(tb + st.c_node[e]) * (K * K) + x * K + p];
    return s;
}

struct DownBytes {
    Bytes b;
    uint8_t kind;
};

template <int K, int CPT, bool GENERAL>
__global__ void __launch_bounds__(CPT == 1 ? 768 : (CPT == 2 ? 512 : 256)) down_kernel(const WalkArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LROW = K + 2, KK = K * K;
    const TileDesc tile = a.tiles[blockIdx.x];
    const int tcpad = a.ncg * CPT;
    Smem<K> sm(smem_raw, a.ncat_max, tcpad);
    const int tid = threadIdx.x, p = tid % K, cg = tid / K;
    const size_t stage_elems = (size_t)a.ncat_max * K * LROW;
    const int n_steps = a.n_steps;

    int lc[CPT], col[CPT], tcat[CPT], ocol[CPT];
    size_t tb[CPT];
    bool act[CPT];
#pragma unroll
    for (int j = 0; j < CPT; j++) {
        lc[j] = cg * CPT + j;
        act[j] = (cg < a.ncg) && (lc[j] < tile.ncols);
        col[j] = tile.col0 + (act[j] ? lc[j] : 0);
        tcat[j] = a.col_cat[col[j]] - tile.cat0;
        ocol[j] = a.orig_col[col[j]];
        tb[j] = (size_t)(tile.cat0 + tcat[j]) * a.n_nodes;
    }
    const double prior_p = a.prior[p];

    prefetch_prog(a.down, n_steps, 0, sm.prog);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    // ring: step i uses stage i % NSTAGE; the table of step i + NSTAGE is requested right after
    // the barrier of step i, when nobody reads stage i % NSTAGE any more.
    for (int s = 0; s < NSTAGE; s++) {
        if (s < n_steps) prefetch_table<K>(a.T_row, a.n_nodes, tile, sm.prog[s].node, sm.tab + s * stage_elems);
        cp_async_commit();
    }
    auto step_ref = [&](int i) -> const Step & { return sm.prog[((i / PCHUNK) & 1) * PCHUNK + (i % PCHUNK)]; };
    auto load_dbytes = [&](const Step &s, int j) -> DownBytes {
        DownBytes d;
        d.b = load_bytes<GENERAL>(a, s, ocol[j], true);
        d.kind = GENERAL ? a.kindm[(size_t)s.ent * a.ncols + col[j]] : (s.parent < 0 ? KIND_ROOT : KIND_MSG);
        return d;
    };
    // values of the first two children (observed: table entry; internal: stored up-message) and the root-like base
    auto load_looks = [&](const Step &s, const DownBytes &d, int j) -> Looks {
        Looks l;
        l.base = prior_p; l.g[0] = l.g[1] = 1.0;
        if (d.kind == KIND_ROOT && d.b.op != BNK_NONE) l.base = a.T_col[(tb[j] + s.node) * KK + p * K + d.b.op];
        if (d.kind == KIND_MSG || d.kind == KIND_ROOT) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                if (e >= s.n_child || !d.b.pu[e]) continue;
                if (d.b.ou[e] != BNK_NONE) l.g[e] = a.T_col[(tb[j] + s.c_node[e]) * KK + d.b.ou[e] * K + p];
                else if (s.c_slot[e] == SLOT_WIDE) l.g[e] = a.msg[((size_t)s.c_ent[e] * K + p) * a.ncols + col[j]];
                else if (s.c_ent[e] >= 0) l.g[e] = a.msg[((size_t)s.c_ent[e] * a.ncols + col[j]) * K + p];
            }
        }
        return l;
    };
    DownBytes B0[CPT], B1[CPT];
    Looks L0[CPT];
#pragma unroll
    for (int j = 0; j < CPT; j++) {
        if (act[j] && n_steps > 0) B0[j] = load_dbytes(step_ref(0), j);
        if (act[j] && n_steps > 1) B1[j] = load_dbytes(step_ref(1), j);
        if (act[j] && n_steps > 0) L0[j] = load_looks(step_ref(0), B0[j], j);
    }
    cp_async_wait<NSTAGE - 1>();
    __syncthreads();

    for (int i = 0; i < n_steps; i++) {
        const Step st = step_ref(i);
        double *gs = sm.gs + (size_t)(i & 1) * tcpad * LROW;
        const double *stage = sm.tab + (i % NSTAGE) * stage_elems;
        DownBytes B2[CPT];
        Looks L1[CPT];
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            if (!act[j]) continue;
            if (i + 2 < n_steps) B2[j] = load_dbytes(step_ref(i + 2), j);
            if (i + 1 < n_steps) L1[j] = load_looks(step_ref(i + 1), B1[j], j);
        }
        uint8_t kind[CPT];
        double U[CPT];
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            kind[j] = KIND_NONE; U[j] = 0.0;
            if (!act[j]) continue;
            kind[j] = B0[j].kind;
            if (kind[j] != KIND_MSG && kind[j] != KIND_ROOT) continue;
            const Bytes b = B0[j].b;
            // ---- from-above vector ----
            double D;
            if (kind[j] == KIND_ROOT) {
                D = L0[j].base;
            } else {
                const double2 *row2 = reinterpret_cast<const double2 *>(stage + ((size_t)tcat[j] * K + p) * LROW);
                const double2 *t2 = (st.slot < a.smem_slots)
                    ? reinterpret_cast<const double2 *>(sm.stack + ((size_t)st.slot * tcpad + lc[j]) * K)
                    : reinterpret_cast<const double2 *>(a.spill + ((size_t)(st.slot - a.smem_slots) * a.ncols + col[j]) * K);
                double acc0 = 0.0, acc1 = 0.0;
                int mhi = 0;
#pragma unroll
                for (int x2 = 0; x2 < K / 2; x2++) {
                    const double2 l = row2[x2], t = t2[x2];
                    acc0 = fma(l.x, t.x, acc0);
                    acc1 = fma(l.y, t.y, acc1);
                    mhi = max(mhi, max(__double2hiint(t.x), __double2hiint(t.y)));
                }
                int e;
                const double scale = pow2_scale_from_hi(mhi, e);
                D = (acc0 + acc1) * scale;
            }
            // ---- children: messages, product, and the from-above vectors they receive.  Child slots
            // never alias this step's own slot (tree_plan.cpp), so they are written without a barrier.
            if (st.n_child <= 2) {
                const double g0 = (st.n_child > 0 && b.pu[0]) ? child_value(a, st, b, L0[j], 0, tb[j], col[j], p, K) : 1.0;
                const double g1 = (st.n_child > 1 && b.pu[1]) ? child_value(a, st, b, L0[j], 1, tb[j], col[j], p, K) : 1.0;
                U[j] = D * (g0 * g1);
                if (st.n_child > 0 && b.pu[0] && b.ou[0] == BNK_NONE && st.c_ent[0] >= 0) {
                    const double t = D * g1;
                    const int sl = st.c_slot[0];
                    if (sl == SLOT_WIDE) a.tmsg[((size_t)st.c_ent[0] * K + p) * a.ncols + col[j]] = t;
                    else if (sl < a.smem_slots) sm.stack[((size_t)sl * tcpad + lc[j]) * K + p] = t;
                    else a.spill[((size_t)(sl - a.smem_slots) * a.ncols + col[j]) * K + p] = t;
                }
                if (st.n_child > 1 && b.pu[1] && b.ou[1] == BNK_NONE && st.c_ent[1] >= 0) {
                    const double t = D * g0;
                    const int sl = st.c_slot[1];
                    if (sl == SLOT_WIDE) a.tmsg[((size_t)st.c_ent[1] * K + p) * a.ncols + col[j]] = t;
                    else if (sl < a.smem_slots) sm.stack[((size_t)sl * tcpad + lc[j]) * K + p] = t;
                    else a.spill[((size_t)(sl - a.smem_slots) * a.ncols + col[j]) * K + p] = t;
                }
            } else {
                // value of child e's message for entry p, and whether the child receives a from-above vector
                auto child_val = [&](int e, bool &push) -> double {
                    const ChildRef ch = a.down_child[st.child_begin + e];
                    push = false;
                    const bool pu = (e == 0) ? (b.pu[0] != 0) : (e == 1) ? (b.pu[1] != 0)
                                  : (!(GENERAL && a.present) || a.present[(size_t)ch.node * a.ncols + ocol[j]] != 0);
                    if (!pu) return 1.0;
                    const uint8_t ou = (e == 0) ? b.ou[0] : (e == 1) ? b.ou[1]
                                     : ((GENERAL || ch.ent < 0) ? a.obs[(size_t)ch.node * a.ncols + ocol[j]] : (uint8_t)BNK_NONE);
                    push = (ou == BNK_NONE && ch.ent >= 0);
                    if (e == 0) return child_value(a, st, b, L0[j], 0, tb[j], col[j], p, K);
                    if (e == 1) return child_value(a, st, b, L0[j], 1, tb[j], col[j], p, K);
                    if (ou != BNK_NONE) return a.T_col[(tb[j] + ch.node) * KK + ou * K + p];
                    if (ch.ent < 0) {
                        double s = 0.0;
                        for (int x = 0; x < K; x++) s += a.T_col[(tb[j] + ch.node) * KK + x * K + p];
                        return s;
                    }
                    if (ch.slot == SLOT_WIDE) return a.msg[((size_t)ch.ent * K + p) * a.ncols + col[j]];
                    return a.msg[((size_t)ch.ent * a.ncols + col[j]) * K + p];
                };
                auto push_t = [&](int e, double t) {
                    const ChildRef chp = a.down_child[st.child_begin + e];
                    const int sl = chp.slot;
                    if (sl == SLOT_WIDE) a.tmsg[((size_t)chp.ent * K + p) * a.ncols + col[j]] = t;
                    else if (sl < a.smem_slots) sm.stack[((size_t)sl * tcpad + lc[j]) * K + p] = t;
                    else a.spill[((size_t)(sl - a.smem_slots) * a.ncols + col[j]) * K + p] = t;
                };
                if (st.n_child <= MAXDEG_LOCAL) {  // prefix / suffix products in registers
                    double gv[MAXDEG_LOCAL], suf[MAXDEG_LOCAL];
                    bool push[MAXDEG_LOCAL];
                    const int nch = st.n_child;
                    double G = 1.0;
                    for (int e = 0; e < nch; e++) { gv[e] = child_val(e, push[e]); G *= gv[e]; }
                    U[j] = D * G;
                    {
                        double s = 1.0;
                        for (int e = nch - 1; e >= 0; e--) { suf[e] = s; s *= gv[e]; }
                    }
                    double pre = D;
                    for (int e = 0; e < nch; e++) {
                        if (push[e]) push_t(e, pre * suf[e]);
                        pre *= gv[e];
                    }
                } else {  // any arity (Factorize.getProduct multiplies any number of factors): values re-read per child
                    double G = 1.0;
                    for (int e = 0; e < st.n_child; e++) { bool pe; G *= child_val(e, pe); }
                    U[j] = D * G;
                    for (int e = 0; e < st.n_child; e++) {
                        bool pe, dummy;
                        (void)child_val(e, pe);
                        if (!pe) continue;
                        double t = D;
                        for (int w = 0; w < st.n_child; w++)
                            if (w != e) t *= child_val(w, dummy);
                        push_t(e, t);
                    }
                }
            }
            gs[(size_t)lc[j] * LROW + p] = U[j];
        }
        cp_async_wait<NSTAGE - 2>();  // table of step i + 1 has landed (this thread's copies)
        __syncthreads();              // ... and everybody's; gs and the child slots are visible
        if (i + NSTAGE < n_steps)
            prefetch_table<K>(a.T_row, a.n_nodes, tile, step_ref(i + NSTAGE).node, sm.tab + (i % NSTAGE) * stage_elems);
        if (i % PCHUNK == 0) prefetch_prog(a.down, n_steps, i / PCHUNK + 1, sm.prog + (((i / PCHUNK) + 1) & 1) * PCHUNK);
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
#pragma unroll
        for (int j = 0; j < CPT; j++) { B0[j] = B1[j]; B1[j] = B2[j]; L0[j] = L1[j]; }
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
