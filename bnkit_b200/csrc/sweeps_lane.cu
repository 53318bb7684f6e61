// sweeps_lane.cu -- lane walker: the three tree sweeps for deep trees (and the narrow top of wide ones), K = 20.
//
// The column-owned walk (tree_plan.cpp) is one long dependent chain per column: a 10,000-level caterpillar is
// 10,000 steps that no amount of columns can overlap.  What a step costs is therefore latency, and the r1 walkers
// (sweeps_warp.cu: 5 lanes share a column, non-uniform table reads) spent ~2,000 cycles on it, bound by
// shared-memory wavefronts: every warp re-read the whole 20x20 table for 6 columns.  Here
//   * a THREAD owns a column (32 columns per CTA) and keeps the whole 20-state vector it contracts in registers,
//     exactly like the level kernels of wide.cu -- so table rows are read as warp-uniform 128-bit broadcasts
//     (one wavefront feeds 32 columns x 2 entries) and masks / evidence are per-thread predicates;
//   * the 20 OUTPUTS of a step are split over the CTA's four warps (five each): four short dependency chains
//     run on the four schedulers of the SM instead of one long one, and the new message is handed from warp to
//     warp through the shared-memory stack slot it has to be written to anyway -- one named barrier per step
//     (tree_plan.cpp never lets a step write a slot it reads);
//   * everything a step needs arrives through a ring of shared-memory stages filled by cp.async (LDGSTS) several
//     steps ahead by the same threads: the walk program (32-step chunks), the node's table, the tables of its
//     childless children, their observed states (32 contiguous bytes per node and tile, from a tiled copy of
//     the observation matrix made once per call) and, for children handled elsewhere, their messages
//     ([ent][state][column] rows: 256 contiguous bytes per state).  No producer warps, no mbarriers.
// Arithmetic, association order and scan order are those of the other kernels (and of the oracle): joint adds
// in child order then L_v[p][x] + G[x], first strict maximum scanning x ascending (argmax_row, wide_common.cuh);
// sum-product in linear space with an exact 2^-e rescale per (node, column).
// Preconditions (host-checked): evidence on childless nodes only, at most two children per node, one rate category
// per tile.  MASKED = true takes a presence mask (per-column forests, Prediction.getTree src/asr/Prediction.java:329-402):
// the presence bytes of a step's node, its parent and its two children ride in the stage next to the observed
// states, and every thread classifies its column -- absent / not connected (KIND_NONE), component root (KIND_ROOT:
// the prior joins, the log-likelihood or the root state closes the component) or message (KIND_MSG) -- exactly as
// the general walker of sweeps.cu does; absent children contribute the unit vector.
#include "wide_common.cuh"

#include <cstddef>

namespace bnk {

namespace {
constexpr int LK = 20;
constexpr int LCOLS = 32;              // columns per CTA = lanes
constexpr int LWARPS = 4;              // consumer warps per CTA
constexpr int LPW = LK / LWARPS;       // outputs per warp
constexpr int LTHREADS = LWARPS * 32;
constexpr int LD = 4;                  // stages in flight ahead of the step being computed
constexpr int LNS = LD + 1;            // ring stages
constexpr int LPCH = 32;               // walk-program steps per chunk
constexpr int LPBUF = 3;               // program chunk buffers
constexpr int L_PROG = LPBUF * LPCH * (int)sizeof(Step);
constexpr int L_OBS = 256;             // a stage starts with 8 rows of 32 state / presence bytes
constexpr int L_TAB = LK * LK * 8;     // one table
constexpr int L_CH = LK * LCOLS * 8;   // area of one child: its table (childless) or a [state][column] message tile
constexpr int L_STAGE = L_OBS + L_TAB + 2 * L_CH;
constexpr int L_SLOT = LK * LCOLS * 8; // one stack slot: [state][column]
constexpr int L_UROW = LCOLS + 1;      // posterior tile rows are padded (transposed reads)
constexpr int L_UT = LK * L_UROW * 8;
constexpr int L_MISC = 512;            // prior / log prior (160 B) + original column of every lane (128 B)

__device__ __forceinline__ void cp16(unsigned dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp8(unsigned dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cta_bar() { asm volatile("bar.sync 0;\n" ::: "memory"); }

__device__ __forceinline__ double l_pow2_scale(int mhi, int &e)
{
    const int be = mhi >> 20;
    if (be == 0) { e = 0; return 1.0; }
    e = be - 1023;
    return __hiloint2double((2046 - be) << 20, 0);
}

#define LS_I32(off) (*reinterpret_cast<const int32_t *>(smem + (off)))
#define LS_F64(off) (*reinterpret_cast<double *>(smem + (off)))
#define LREC(po, member) LS_I32((po) + (int)offsetof(Step, member))

__device__ __forceinline__ int rec_off(int j) { return ((j / LPCH) % LPBUF) * (LPCH * (int)sizeof(Step)) + (j % LPCH) * (int)sizeof(Step); }

// copy one K x K table (3,200 bytes = 200 x 16) into shared memory, all threads of the CTA
__device__ __forceinline__ void copy_table(unsigned dst, const double *src, int tid)
{
    const char *s = reinterpret_cast<const char *>(src);
    cp16(dst + tid * 16, s + tid * 16);
    if (tid + LTHREADS < L_TAB / 16) cp16(dst + (tid + LTHREADS) * 16, s + (tid + LTHREADS) * 16);
}

// program chunk c (32 steps = 2,048 bytes = 128 x 16): one 16-byte piece per thread
__device__ __forceinline__ void copy_prog(unsigned sbase, const Step *prog, int n_steps, int c, int tid)
{
    const int first = c * LPCH;
    if (first >= n_steps) return;
    const int bytes = (n_steps - first < LPCH ? n_steps - first : LPCH) * (int)sizeof(Step);
    if (tid * 16 < bytes)
        cp16(sbase + (c % LPBUF) * (LPCH * (int)sizeof(Step)) + tid * 16, reinterpret_cast<const char *>(prog + first) + tid * 16);
}

struct LaneCtx {
    int warp, lane, tid;
    int lc, col, ocol;   // tile column (clamped), sorted column, original column
    bool valid;
    size_t cat_off;      // offset of the tile's category inside the table arrays, in doubles
    int tile_col;        // first byte of this tile inside a row of the tiled observation matrix
};

// Requests everything step j needs into its ring stage (commit is the caller's).  DOWN = false: the up-sweeps
// (node table and child tables from a.T_row; messages of SLOT_WIDE children from a.msg); DOWN = true: the
// down-sweep (node table from a.T_row = P^T, child tables from a.T_aux = P, messages of ALL internal children).
template <bool DOWN, bool MASKED>
__device__ __forceinline__ void issue_step(const WalkArgs &a, const LaneCtx &c, unsigned char *smem, unsigned sbase, int j)
{
    if (j >= a.n_steps) return;
    const int po = rec_off(j);
    const int node = LREC(po, node), parent = LREC(po, parent);
    const unsigned sb = sbase + L_PROG + (j % LNS) * L_STAGE;
    const size_t ncols = (size_t)a.ncols;
    if (parent >= 0) copy_table(sb + L_OBS, a.T_row + c.cat_off + (size_t)(unsigned)node * (LK * LK), c.tid);
    if (MASKED && c.tid >= 8 && c.tid < 16) {
        // presence rows of the stage: 2 = node, 3 = parent, 4 / 5 = children (32 bytes each, two 16-byte pieces)
        const int r = (c.tid - 8) >> 1, half = (c.tid - 8) & 1;
        const int who = r == 0 ? node : (r == 1 ? parent : (r == 2 ? LREC(po, c_node[0]) : LREC(po, c_node[1])));
        if (who >= 0) cp16(sb + (2 + r) * 32 + half * 16, a.present_t + (size_t)(unsigned)who * a.tile_stride + c.tile_col + half * 16);
    }
    const double *ctab = DOWN ? a.T_aux : a.T_row;
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const int cn = e == 0 ? LREC(po, c_node[0]) : LREC(po, c_node[1]);
        const int ce = e == 0 ? LREC(po, c_ent[0]) : LREC(po, c_ent[1]);
        const int cs = e == 0 ? LREC(po, c_slot[0]) : LREC(po, c_slot[1]);
        if (cn < 0) continue;
        const unsigned area = sb + L_OBS + L_TAB + e * L_CH;
        if (ce < 0) {  // childless child: its table and the 32 observed states of the tile
            copy_table(area, ctab + c.cat_off + (size_t)(unsigned)cn * (LK * LK), c.tid);
            if (c.tid >= 4 * e && c.tid < 4 * e + 2)
                cp16(sb + e * 32 + (c.tid - 4 * e) * 16, a.obs_t + (size_t)(unsigned)cn * a.tile_stride + c.tile_col + (c.tid - 4 * e) * 16);
        } else if (DOWN || cs == SLOT_WIDE) {  // a message stored in HBM as [ent][state][column]: this warp's five rows
            const double *m = a.msg + (size_t)(unsigned)ce * (LK * ncols) + c.col;
#pragma unroll
            for (int k = 0; k < LPW; k++) {
                const int x = c.warp * LPW + k;
                cp8(area + x * (LCOLS * 8) + c.lane * 8, m + (size_t)x * ncols);
            }
        }
    }
}

__device__ __forceinline__ LaneCtx lane_ctx(const WalkArgs &a, const TileDesc &tile)
{
    LaneCtx c;
    c.tid = threadIdx.x; c.warp = threadIdx.x >> 5; c.lane = threadIdx.x & 31;
    c.valid = c.lane < tile.ncols;
    c.lc = c.valid ? c.lane : tile.ncols - 1;  // spare lanes shadow the tile's last column
    c.col = tile.col0 + c.lc;
    c.ocol = a.orig_col[c.col];
    c.cat_off = (size_t)tile.cat0 * a.n_nodes * (LK * LK);
    c.tile_col = (a.tile_base + (int)blockIdx.x) * LCOLS;
    return c;
}

// the 20-vector a child contributes to its parent in an up-sweep, for this thread's column
template <int MODE>
__device__ __forceinline__ void child_up_vector(unsigned char *smem, int so, int stack0, int e, int cn, int cs, int lane,
                                                double (&v)[LK], bool &isob)
{
    isob = false;
    if (cs >= 0) {  // left on the stack by an earlier step
        const int off = stack0 + cs * L_SLOT + lane * 8;
#pragma unroll
        for (int x = 0; x < LK; x++) v[x] = LS_F64(off + x * (LCOLS * 8));
    } else if (cs == SLOT_WIDE) {  // handled by a level kernel: the message tile of this stage
        const int off = so + L_OBS + L_TAB + e * L_CH + lane * 8;
#pragma unroll
        for (int x = 0; x < LK; x++) v[x] = LS_F64(off + x * (LCOLS * 8));
    } else {  // childless: column ob of its row-major table, T[x][ob]
        const int tb = so + L_OBS + L_TAB + e * L_CH;
        const int ob = smem[so + e * 32 + lane];
        if (ob != BNK_NONE) {
            isob = true;
#pragma unroll
            for (int x = 0; x < LK; x++) v[x] = LS_F64(tb + (x * LK + ob) * 8);
        } else {  // uninstantiated childless node: maximise / marginalise its own state
#pragma unroll
            for (int x = 0; x < LK; x++) {  // unrolled: v[] must stay in registers
                double r = LS_F64(tb + (x * LK) * 8);
                for (int y = 1; y < LK; y++) {
                    const double t = LS_F64(tb + (x * LK + y) * 8);
                    r = MODE == 0 ? (t > r ? t : r) : r + t;
                }
                v[x] = r;
            }
        }
    }
    (void)cn;
}
}  // namespace

size_t lane_walk_smem_bytes(int slots, bool down)
{
    return (size_t)L_PROG + (size_t)LNS * L_STAGE + (size_t)(slots + 1) * L_SLOT + (down ? 2 * (size_t)L_UT : 0) + L_MISC;
}
int lane_walk_cols() { return LCOLS; }

// ------------------------------------------------------------------------------------
// up-sweeps.  MODE 0: max-product in log space (joint), MODE 1: sum-product, linear, rescaled.
// ------------------------------------------------------------------------------------
template <int MODE, bool MASKED>
__global__ void __launch_bounds__(LTHREADS, 2) lane_up_kernel(const WalkArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    const TileDesc tile = a.tiles[blockIdx.x];
    const LaneCtx c = lane_ctx(a, tile);
    const int n_steps = a.n_steps;
    const size_t ncols = (size_t)a.ncols;
    const int stack0 = L_PROG + LNS * L_STAGE;
    const int dummy = stack0 + a.smem_slots * L_SLOT;
    const int misc = dummy + L_SLOT;
    if (c.tid < LK) LS_F64(misc + c.tid * 8) = a.prior[c.tid];
    if (n_steps <= 0) return;

    // prologue: program chunk 0, then the stages of the first LD steps (one commit group per step)
    copy_prog(sbase, a.up, n_steps, 0, c.tid);
    cp_commit();
    cp_wait<0>();
    __syncthreads();
#pragma unroll
    for (int j = 0; j < LD; j++) {
        if (j == 0) copy_prog(sbase, a.up, n_steps, 1, c.tid);
        issue_step<false, MASKED>(a, c, smem, sbase, j);
        cp_commit();
    }
    int esum = 0;
    double logacc = 0.0;

    for (int i = 0; i < n_steps; i++) {
        cp_wait<LD - 1>();  // this thread's copies for step i have landed ...
        cta_bar();          // ... and everybody's; the message written by step i - 1 is visible
        if ((i % LPCH) == 0 && i > 0) copy_prog(sbase, a.up, n_steps, i / LPCH + 1, c.tid);
        issue_step<false, MASKED>(a, c, smem, sbase, i + LD);
        cp_commit();

        const int po = rec_off(i);
        const int ent = LREC(po, ent), slot = LREC(po, slot);
        const int so = L_PROG + (i % LNS) * L_STAGE;
        int cn0 = LREC(po, c_node[0]), cn1 = LREC(po, c_node[1]);
        const int cs0 = LREC(po, c_slot[0]), cs1 = LREC(po, c_slot[1]);
        // this column's case: root-like when the parent is absent (or there is none); children that are absent in
        // this column vanish (cn < 0 from here on); KIND_NONE = the node is absent or has no BN variable
        bool rootstep = LREC(po, parent) < 0;
        int kind = rootstep ? KIND_ROOT : KIND_MSG;
        if (MASKED) {
            const bool pv = smem[so + 2 * 32 + c.lane] != 0;
            if (!rootstep && smem[so + 3 * 32 + c.lane] == 0) rootstep = true;
            if (cn0 >= 0 && smem[so + 4 * 32 + c.lane] == 0) cn0 = -1;
            if (cn1 >= 0 && smem[so + 5 * 32 + c.lane] == 0) cn1 = -1;
            kind = !pv ? KIND_NONE : (rootstep ? ((cn0 < 0 && cn1 < 0) ? KIND_NONE : KIND_ROOT) : KIND_MSG);
            if (c.warp == 0 && c.valid) {
                if (MODE == 0) a.kind[(size_t)(unsigned)ent * ncols + c.col] = (uint8_t)kind;
                else if (a.kindm) a.kindm[(size_t)(unsigned)ent * ncols + c.col] = (uint8_t)kind;
            }
        }

        // ---- fold the (at most two) children: G lives in registers, identical in the four warps ----
        double G[LK];
        {
            double v0[LK], v1[LK];
            bool ob0 = false, ob1 = false;
            if (cn0 >= 0) child_up_vector<MODE>(smem, so, stack0, 0, cn0, cs0, c.lane, v0, ob0);
            if (cn1 >= 0) child_up_vector<MODE>(smem, so, stack0, 1, cn1, cs1, c.lane, v1, ob1);
            // a root combines [prior, observed children, unobserved children] (the order the factors sit in the
            // reference's bucket); with one child of each kind that fixes the association
            const bool swap = MODE == 0 && rootstep && cn0 >= 0 && cn1 >= 0 && !ob0 && ob1;
#pragma unroll
            for (int x = 0; x < LK; x++) {
                const double a0 = cn0 >= 0 ? v0[x] : (MODE == 0 ? 0.0 : 1.0);
                const double a1 = cn1 >= 0 ? v1[x] : (MODE == 0 ? 0.0 : 1.0);
                const double c0 = swap ? a1 : a0, c1 = swap ? a0 : a1;
                const double base = rootstep ? LS_F64(misc + x * 8) : (MODE == 0 ? 0.0 : 1.0);
                G[x] = MODE == 0 ? (base + c0) + c1 : (base * c0) * c1;
            }
        }
        const int sdst = (slot >= 0 ? stack0 + slot * L_SLOT : dummy) + c.lane * 8;
        const int tv = so + L_OBS;
        if (MODE == 0) {
            uint8_t *pr = a.ptr + ((size_t)(unsigned)ent * ncols + c.col) * LK + c.warp * LPW;
            if (kind == KIND_MSG) {
#pragma unroll
                for (int k = 0; k < LPW; k++) {
                    const int p = c.warp * LPW + k;
                    double best;
                    int bi;
                    argmax_row<LK>(reinterpret_cast<const double *>(smem + tv) + p * LK, G, best, bi);
                    LS_F64(sdst + p * (LCOLS * 8)) = best;
                    if (c.valid) pr[k] = (uint8_t)bi;
                }
            } else if (kind == KIND_ROOT) {  // unit table: every output is max_x G[x]
                double b = -CUDART_INF;
                int bx = 0;
#pragma unroll
                for (int x = 0; x < LK; x++) {
                    const double v = 0.0 + G[x];
                    if (v > b) { b = v; bx = x; }
                }
#pragma unroll
                for (int k = 0; k < LPW; k++) {
                    LS_F64(sdst + (c.warp * LPW + k) * (LCOLS * 8)) = b;
                    if (c.valid) pr[k] = (uint8_t)bx;
                }
            }
        } else if (kind != KIND_NONE) {
            int mhi = 0;
#pragma unroll
            for (int x = 0; x < LK; x++) mhi = max(mhi, __double2hiint(G[x]));
            int e;
            const double scale = l_pow2_scale(mhi, e);
            esum += e;
            if (kind == KIND_MSG) {
#pragma unroll
                for (int k = 0; k < LPW; k++) {
                    const int p = c.warp * LPW + k;
                    const double2 *row2 = reinterpret_cast<const double2 *>(smem + tv + p * (LK * 8));
                    double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
                    for (int x2 = 0; x2 < LK / 2; x2++) {
                        const double2 l = row2[x2];
                        acc0 = fma(l.x, G[2 * x2], acc0);
                        acc1 = fma(l.y, G[2 * x2 + 1], acc1);
                    }
                    const double m = (acc0 + acc1) * scale;
                    LS_F64(sdst + p * (LCOLS * 8)) = m;
                    if (a.msg && c.valid) a.msg[((size_t)(unsigned)ent * LK + p) * ncols + c.col] = m;
                }
            } else {
                double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
                for (int x2 = 0; x2 < LK / 2; x2++) { acc0 += G[2 * x2]; acc1 += G[2 * x2 + 1]; }
                logacc += log((acc0 + acc1) * scale);
            }
        }
    }
    if (MODE == 1 && a.out_logl && c.warp == 0 && c.valid) {
        const int ew = a.esum_wide ? a.esum_wide[c.col] : 0;
        const double lw = (MASKED && a.logl_wide) ? a.logl_wide[c.col] : 0.0;   // components closed inside the wide levels
        a.out_logl[c.ocol] = logacc + lw + (double)(esum + ew) * 0.693147180559945309417232121458;
    }
}

// ------------------------------------------------------------------------------------
// down-sweep: T (from above, stack) -> D = P^T T -> posterior, and the from-above vectors of internal children
// ------------------------------------------------------------------------------------
template <bool MASKED>
__global__ void __launch_bounds__(LTHREADS, 2) lane_down_kernel(const WalkArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    const TileDesc tile = a.tiles[blockIdx.x];
    const LaneCtx c = lane_ctx(a, tile);
    const int n_steps = a.n_steps;
    const size_t ncols = (size_t)a.ncols;
    const int stack0 = L_PROG + LNS * L_STAGE;
    const int ut0 = stack0 + (a.smem_slots + 1) * L_SLOT;
    const int misc = ut0 + 2 * L_UT;
    if (c.tid < LK) LS_F64(misc + c.tid * 8) = a.prior[c.tid];
    if (c.tid < LCOLS) *reinterpret_cast<int32_t *>(smem + misc + 256 + c.tid * 4) = c.ocol;  // warp 0: lane == column
    if (n_steps <= 0) return;

    copy_prog(sbase, a.down, n_steps, 0, c.tid);
    cp_commit();
    cp_wait<0>();
    __syncthreads();
#pragma unroll
    for (int j = 0; j < LD; j++) {
        if (j == 0) copy_prog(sbase, a.down, n_steps, 1, c.tid);
        issue_step<true, MASKED>(a, c, smem, sbase, j);
        cp_commit();
    }
    // the posterior of step i is normalised and stored after the barrier of step i + 1: warp w then owns tile
    // columns 8 w .. 8 w + 7, four threads per column with five states each (40 contiguous bytes per thread)
    const int ecol = c.warp * 8 + (c.lane >> 2), ex0 = (c.lane & 3) * LPW;
    auto emit = [&](int i_prev, int q_prev) {
        if (q_prev < 0 || !a.out_post) return;
        const int ub = ut0 + (i_prev & 1) * L_UT + ecol * 8;
        double u[LPW], s = 0.0;
#pragma unroll
        for (int k = 0; k < LPW; k++) { u[k] = LS_F64(ub + (ex0 + k) * (L_UROW * 8)); s += u[k]; }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (ecol < tile.ncols) {
            const double r = w_rcp(s);
            const int oc = *reinterpret_cast<const int32_t *>(smem + misc + 256 + ecol * 4);
            double *o = a.out_post + ((size_t)q_prev * ncols + oc) * LK + ex0;
#pragma unroll
            for (int k = 0; k < LPW; k++) __stcs(o + k, u[k] * r);
        }
    };
    int q_prev = -1;

    for (int i = 0; i < n_steps; i++) {
        cp_wait<LD - 1>();
        cta_bar();
        if ((i % LPCH) == 0 && i > 0) copy_prog(sbase, a.down, n_steps, i / LPCH + 1, c.tid);
        issue_step<true, MASKED>(a, c, smem, sbase, i + LD);
        cp_commit();
        if (i > 0) emit(i - 1, q_prev);

        const int po = rec_off(i);
        const int ent = LREC(po, ent), slot = LREC(po, slot);
        const int so = L_PROG + (i % LNS) * L_STAGE;
        q_prev = a.qrow ? a.qrow[ent] : ent;
        // the same classification as in the up-sweep (no evidence on nodes with children): absent children count
        // as the unit vector and receive nothing; a node without a BN variable emits NaN and pushes nothing
        bool rootstep = LREC(po, parent) < 0;
        bool pu[2] = {true, true};
        bool none = false;
        if (MASKED) {
            const bool pv = smem[so + 2 * 32 + c.lane] != 0;
            if (!rootstep && smem[so + 3 * 32 + c.lane] == 0) rootstep = true;
            pu[0] = LREC(po, c_node[0]) >= 0 && smem[so + 4 * 32 + c.lane] != 0;
            pu[1] = LREC(po, c_node[1]) >= 0 && smem[so + 5 * 32 + c.lane] != 0;
            none = !pv || (rootstep && !pu[0] && !pu[1]);
        }

        // ---- D[x] = sum_p P[p][x] T[p] for this warp's five x ----
        double D[LPW];
        if (!rootstep) {
            double T[LK];
            const int off = stack0 + slot * L_SLOT + c.lane * 8;
            int mhi = 0;
#pragma unroll
            for (int x = 0; x < LK; x++) { T[x] = LS_F64(off + x * (LCOLS * 8)); mhi = max(mhi, __double2hiint(T[x])); }
            int e;
            const double scale = l_pow2_scale(mhi, e);
#pragma unroll
            for (int k = 0; k < LPW; k++) {
                const double2 *row2 = reinterpret_cast<const double2 *>(smem + so + L_OBS + (c.warp * LPW + k) * (LK * 8));
                double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
                for (int y2 = 0; y2 < LK / 2; y2++) {
                    const double2 l = row2[y2];
                    acc0 = fma(l.x, T[2 * y2], acc0);
                    acc1 = fma(l.y, T[2 * y2 + 1], acc1);
                }
                D[k] = (acc0 + acc1) * scale;
            }
        } else {
#pragma unroll
            for (int k = 0; k < LPW; k++) D[k] = LS_F64(misc + (c.warp * LPW + k) * 8);
        }
        // ---- the children's messages for those x ----
        double g[2][LPW];
        bool push[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            int cn = e == 0 ? LREC(po, c_node[0]) : LREC(po, c_node[1]);
            const int ce = e == 0 ? LREC(po, c_ent[0]) : LREC(po, c_ent[1]);
            if (MASKED && !pu[e]) cn = -1;
            push[e] = cn >= 0 && ce >= 0 && !none;
            const int area = so + L_OBS + L_TAB + e * L_CH;
            if (cn < 0) {
#pragma unroll
                for (int k = 0; k < LPW; k++) g[e][k] = 1.0;
            } else if (ce >= 0) {  // stored up-message, this warp's rows
#pragma unroll
                for (int k = 0; k < LPW; k++) g[e][k] = LS_F64(area + (c.warp * LPW + k) * (LCOLS * 8) + c.lane * 8);
            } else {
                const int ob = smem[so + e * 32 + c.lane];
                if (ob != BNK_NONE) {
#pragma unroll
                    for (int k = 0; k < LPW; k++) g[e][k] = LS_F64(area + ((c.warp * LPW + k) * LK + ob) * 8);
                } else {  // uninstantiated childless node: its row sums
#pragma unroll
                    for (int k = 0; k < LPW; k++) {
                        double r = 0.0;
                        for (int y = 0; y < LK; y++) r += LS_F64(area + ((c.warp * LPW + k) * LK + y) * 8);
                        g[e][k] = r;
                    }
                }
            }
        }
        // ---- from-above vectors of internal children, unnormalised posterior ----
#pragma unroll
        for (int e = 0; e < 2; e++) {
            if (!push[e]) continue;
            const int cs = e == 0 ? LREC(po, c_slot[0]) : LREC(po, c_slot[1]);
            const int ce = e == 0 ? LREC(po, c_ent[0]) : LREC(po, c_ent[1]);
#pragma unroll
            for (int k = 0; k < LPW; k++) {
                const int x = c.warp * LPW + k;
                const double t = D[k] * g[1 - e][k];
                if (cs == SLOT_WIDE) { if (c.valid) a.tmsg[((size_t)(unsigned)ce * LK + x) * ncols + c.col] = t; }
                else LS_F64(stack0 + cs * L_SLOT + x * (LCOLS * 8) + c.lane * 8) = t;
            }
        }
        const int ub = ut0 + (i & 1) * L_UT + c.lane * 8;
#pragma unroll
        for (int k = 0; k < LPW; k++)
            LS_F64(ub + (c.warp * LPW + k) * (L_UROW * 8)) = (MASKED && none) ? CUDART_NAN : D[k] * (g[0][k] * g[1][k]);
    }
    cta_bar();
    emit(n_steps - 1, q_prev);
}

// ------------------------------------------------------------------------------------
// tiled copy of a [node][column] byte matrix: out[node][tile * 32 + lane] = in[node][orig_col[tile.col0 + lane]],
// spare lanes repeating the tile's last column -- 32 contiguous, 32-byte aligned bytes per (node, tile)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tile_bytes_kernel(const uint8_t *__restrict__ in, const int32_t *__restrict__ orig_col,
                                                         const TileDesc *__restrict__ tiles, int n_tiles, int tile_base, int n_nodes,
                                                         int ncols, int tile_stride, uint8_t *__restrict__ out)
{
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= n_tiles) return;
    const TileDesc td = tiles[t];
    const int oc = orig_col[td.col0 + (lane < td.ncols ? lane : td.ncols - 1)];
    for (int v = blockIdx.y; v < n_nodes; v += gridDim.y)
        out[(size_t)v * tile_stride + (size_t)(tile_base + t) * LCOLS + lane] = in[(size_t)v * ncols + oc];
}

cudaError_t launch_tile_bytes(const uint8_t *in, const int32_t *orig_col, const TileDesc *tiles, int n_tiles, int tile_base,
                              int n_nodes, int ncols, int tile_stride, uint8_t *out, cudaStream_t st)
{
    if (n_tiles == 0 || n_nodes == 0) return cudaSuccess;
    dim3 grid((n_tiles + 7) / 8, n_nodes < 2048 ? n_nodes : 2048);
    tile_bytes_kernel<<<grid, 256, 0, st>>>(in, orig_col, tiles, n_tiles, tile_base, n_nodes, ncols, tile_stride, out);
    return cudaGetLastError();
}

template <typename KernelT>
static cudaError_t launch_lane(KernelT kern, const WalkArgs &a, int n_tiles, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    kern<<<n_tiles, LTHREADS, smem, st>>>(a);
    return cudaGetLastError();
}

// a.present_t != nullptr selects the masked kernels
cudaError_t launch_lane_up(int mode, const WalkArgs &a, int n_tiles, cudaStream_t st)
{
    if (n_tiles == 0) return cudaSuccess;
    const size_t smem = lane_walk_smem_bytes(a.smem_slots, false);
    if (a.present_t)
        return mode == 0 ? launch_lane(lane_up_kernel<0, true>, a, n_tiles, smem, st) : launch_lane(lane_up_kernel<1, true>, a, n_tiles, smem, st);
    return mode == 0 ? launch_lane(lane_up_kernel<0, false>, a, n_tiles, smem, st) : launch_lane(lane_up_kernel<1, false>, a, n_tiles, smem, st);
}

cudaError_t launch_lane_down(const WalkArgs &a, int n_tiles, cudaStream_t st)
{
    if (n_tiles == 0) return cudaSuccess;
    const size_t smem = lane_walk_smem_bytes(a.smem_slots, true);
    return a.present_t ? launch_lane(lane_down_kernel<true>, a, n_tiles, smem, st) : launch_lane(lane_down_kernel<false>, a, n_tiles, smem, st);
}

}  // namespace bnk
