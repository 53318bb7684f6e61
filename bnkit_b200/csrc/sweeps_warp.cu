// sweeps_warp.cu -- warp-tiled walker: the three tree sweeps for deep trees (and narrow tops), K = 20.
//
// Same walk programs, arithmetic and association order as sweeps_fast.cu; what changes is who owns what.
// There a CTA owned a tile of columns, 20 threads shared a column (one output each) and every step of
// the walk ended in a CTA barrier; on a 10,000-level caterpillar that cost ~400 issued instructions per
// warp and step at 40 % issue utilisation (profiles/r1_v12_bench_caterpillar.json: 1.4 us per step).
// Here
//   * a WARP owns 6 x C columns; 5 lanes share a column and each computes 4 of the 20 outputs for C columns
//     (C = 1..3 columns per lane, register-tiled), so one 128-bit shared-memory read of the node's table
//     feeds 2 x C multiply-adds (or add/compare/selects).  Shared-memory wavefronts are what bounds a walk
//     (profiles/r1_v13_*: a non-uniform LDS.128 costs four wavefronts, one per quarter-warp, however many
//     lanes share an address), so what matters is bytes fetched per column: lane q owns outputs
//     {2q, 2q+1, 10+2q, 11+2q}, which makes the five lanes of a column read 80 contiguous bytes of a table
//     row per load (no bank conflict), and the table traffic is shared by the C columns of a lane;
//   * warps never meet at a CTA barrier: the node's table, the tables of its childless children and (going
//     down) the stored up-messages of its internal children arrive in a ring of shared-memory stages filled
//     by a producer warp with TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx), one full / empty
//     mbarrier pair per stage; the observed states of childless children travel in the same stage (the
//     producer requests them XPD steps ahead), so the consumers' loop contains no global load at all
//     unless a child was handled by a wide level kernel;
//   * the walk program itself streams through a second small ring (32-step chunks, bulk copies).
// Requirements (the host falls back to sweeps_fast.cu / sweeps.cu otherwise): no presence mask, evidence
// on childless nodes only, at most two children per node, one rate category per tile, whole message stack
// in shared memory.
#include "device_common.cuh"

#include <cstddef>

namespace bnk {

namespace {
constexpr int XK = 20;
constexpr int XCW = 6;             // columns per warp (5 lanes each; lanes 30, 31 shadow lanes 25, 26)
constexpr int XPD = 8;             // producer: observed states are requested this many steps ahead
constexpr int XPB = 4;             // walk-program ring: chunks of 32 steps
constexpr int XTAB = XK * XK * 8;  // one table, bytes
constexpr int XPROG = 256;         // byte offsets inside dynamic shared memory: barriers first, ...
constexpr int XSTAGE0 = XPROG + XPB * 32 * (int)sizeof(Step);
constexpr int XMAXNS = 8, XMAXW = 7;

constexpr int XOBS = 128;          // a stage starts with the observed states: 2 children x 64 tile columns
constexpr int XOBS_M = 192;        // MASKED: + one case byte per tile column (kind | child 0 present << 2 | child 1 present << 3)
// area of one child inside a stage: its table (childless) or, going down, the stored up-messages of the tile
__host__ __device__ inline int x_child_bytes(int tile_cols, bool down) { return (down && tile_cols * XK * 8 > XTAB) ? tile_cols * XK * 8 : XTAB; }
// stage = [observed states (+ case bytes)][table of the node][area of child 0][area of child 1]
__host__ __device__ inline int x_stage_bytes(int chb, bool masked = false) { return (masked ? XOBS_M : XOBS) + XTAB + 2 * chb; }
// MASKED: the unit table (zeros in log space, ones in linear space) sits between the program ring and stage 0
__host__ __device__ inline int x_stage0(bool masked) { return XSTAGE0 + (masked ? XTAB : 0); }
// a warp's private area: two exchange buffers and the message stack (+ 1 dummy slot), all "planar" vectors:
// entries 0..9 of column c at c * 80, entries 10..19 at plane + c * 80 (plane = columns * 80 bytes), so that
// consecutive lanes touch consecutive 16-byte chunks and no quarter-warp sees a bank twice
__host__ __device__ inline int x_warp_bytes(int C, int slots) { return (2 + slots + 1) * XCW * C * XK * 8; }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    // test_wait first: in steady state the phase has completed long before anybody asks (the ring runs several
    // steps ahead), and the non-blocking probe returns at once where try_wait may park the warp for a while
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
// non-blocking probe; true = the phase with this parity has completed (acquire, like a successful wait)
__device__ __forceinline__ unsigned mbar_test(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ double x_pow2_scale(int mhi, int &e)
{
    const int be = mhi >> 20;
    if (be == 0) { e = 0; return 1.0; }
    e = be - 1023;
    return __hiloint2double((2046 - be) << 20, 0);
}
__device__ __forceinline__ double x_rcp(double s)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
    r = fma(r, fma(-s, r, 1.0), r);
    r = fma(r, fma(-s, r, 1.0), r);
    return r;
}

#define XS_I32(off) (*reinterpret_cast<const int32_t *>(smem_raw + (off)))
#define XS_D2(off) (*reinterpret_cast<double2 *>(smem_raw + (off)))
#define XSTEP(po, member) XS_I32((po) + (int)offsetof(Step, member))

struct XGeom {
    int chb, stage_bytes;
    unsigned bar_full, bar_empty, bar_prog;
};

// barriers + role split shared by the kernels
__device__ __forceinline__ XGeom x_setup(unsigned sbase, int W, int ns, int chb, bool masked = false)
{
    XGeom g;
    g.chb = chb;
    g.stage_bytes = x_stage_bytes(chb, masked);
    g.bar_full = sbase;
    g.bar_empty = sbase + 8 * XMAXNS;
    g.bar_prog = sbase + 16 * XMAXNS;
    if (threadIdx.x == 0) {
        for (int s = 0; s < ns; s++) {
            mbar_init(g.bar_full + 8 * s, 33);  // the 32 lanes of the obs warp + the issuing lane of the TMA warp
            mbar_init(g.bar_empty + 8 * s, W);  // lane 0 of every consumer warp
        }
        for (int b = 0; b < XPB; b++) mbar_init(g.bar_prog + 8 * b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
    return g;
}

// Two producer warps feed the ring (one alone spent ~860 cycles per step on its own instructions -- as long as
// a consumer step -- and the two kept blocking each other, profiles/r1_v19):
//   warp W      issues the TMA bulk copies of a step (one elected lane): the node's table, the tables of its
//               childless children and, going down, the stored up-messages of its walker-internal children;
//               DOWN = false: tables from T_col; DOWN = true: node table from T_aux, child tables from T_col;
//               it also streams the walk program (32-step chunks);
//   warp W + 1  stores the observed states of the childless children, which it requested XPD steps earlier.
// A stage is full when the 32 lanes of the second and the issuing lane of the first have arrived and the bytes
// have landed.
template <bool DOWN, bool MASKED = false>
__device__ __forceinline__ void x_producer_tma(const WalkArgs &a, const TileDesc &tile, const XGeom &g, unsigned char *smem_raw,
                                               unsigned sbase, int ns, int lane)
{
    constexpr int XOBS = MASKED ? XOBS_M : bnk::XOBS;
    constexpr int XSTAGE0 = bnk::XSTAGE0 + (MASKED ? XTAB : 0);
    const Step *prog = DOWN ? a.down : a.up;
    const int n_steps = a.n_steps;
    if (n_steps <= 0) return;  // nothing to stream (and nobody would complete the first program chunk)
    const size_t ncols = (size_t)a.ncols;
    const size_t cat_off = (size_t)tile.cat0 * a.n_nodes * (XK * XK);
    const double *tab_node = (DOWN ? a.T_aux : a.T_col) + cat_off;
    const double *tab_child = a.T_col + cat_off;
    const unsigned msg_bytes = (unsigned)tile.ncols * XK * 8;
    auto issue_chunk = [&](int j) {
        if (lane == 0 && j * 32 < n_steps) {
            const int cnt = n_steps - j * 32 < 32 ? n_steps - j * 32 : 32;
            const unsigned bar = g.bar_prog + (j & (XPB - 1)) * 8;
            mbar_arrive_tx(bar, cnt * (unsigned)sizeof(Step));
            bulk_g2s(sbase + XPROG + (j & (XPB - 1)) * 32 * (int)sizeof(Step), prog + (size_t)j * 32, cnt * (unsigned)sizeof(Step), bar);
        }
    };
    issue_chunk(0);
    issue_chunk(1);
    int s = 0, ph = 0;
    for (int i = 0; i < n_steps; i++) {
        if ((i & 31) == 0) {
            issue_chunk((i >> 5) + 2);
            mbar_wait(g.bar_prog + ((i >> 5) & (XPB - 1)) * 8, (i >> 7) & 1);
        }
        if (i >= ns) mbar_wait(g.bar_empty + s * 8, ph ^ 1);
        if (lane == 0) {
            const int po = XPROG + (i & (XPB * 32 - 1)) * (int)sizeof(Step);
            const int node = XSTEP(po, node), parent = XSTEP(po, parent);
            const int cn0 = XSTEP(po, c_node[0]), cn1 = XSTEP(po, c_node[1]);
            const int ce0 = XSTEP(po, c_ent[0]), ce1 = XSTEP(po, c_ent[1]);
            const int cs0 = XSTEP(po, c_slot[0]), cs1 = XSTEP(po, c_slot[1]);
            const int so = XSTAGE0 + s * g.stage_bytes;
            const unsigned bar = g.bar_full + s * 8;
            const bool tn = parent >= 0;  // a root contracts with the unit table / uses the prior
            const bool t0 = cn0 >= 0 && ce0 < 0, t1 = cn1 >= 0 && ce1 < 0;
            const bool m0 = DOWN && ce0 >= 0 && cs0 != SLOT_WIDE, m1 = DOWN && ce1 >= 0 && cs1 != SLOT_WIDE;
            const unsigned bytes = (tn ? XTAB : 0) + (t0 ? XTAB : 0) + (t1 ? XTAB : 0) + (m0 ? msg_bytes : 0) + (m1 ? msg_bytes : 0);
            if (bytes) mbar_arrive_tx(bar, bytes); else mbar_arrive(bar);
            const unsigned stg = sbase + so;
            if (tn) bulk_g2s(stg + XOBS, tab_node + (size_t)(unsigned)node * (XK * XK), XTAB, bar);
            if (t0) bulk_g2s(stg + XOBS + XTAB, tab_child + (size_t)(unsigned)cn0 * (XK * XK), XTAB, bar);
            if (t1) bulk_g2s(stg + XOBS + XTAB + g.chb, tab_child + (size_t)(unsigned)cn1 * (XK * XK), XTAB, bar);
            if (m0) bulk_g2s(stg + XOBS + XTAB, a.msg + ((size_t)(unsigned)ce0 * ncols + tile.col0) * XK, msg_bytes, bar);
            if (m1) bulk_g2s(stg + XOBS + XTAB + g.chb, a.msg + ((size_t)(unsigned)ce1 * ncols + tile.col0) * XK, msg_bytes, bar);
        }
        __syncwarp();
        if (++s == ns) { s = 0; ph ^= 1; }
    }
}

template <bool DOWN, bool MASKED = false>
__device__ __forceinline__ void x_producer_obs(const WalkArgs &a, const TileDesc &tile, const XGeom &g, unsigned char *smem_raw,
                                               int ns, int lane)
{
    constexpr int XSTAGE0 = bnk::XSTAGE0 + (MASKED ? XTAB : 0);
    const int n_steps = a.n_steps;
    if (n_steps <= 0) return;
    const size_t ncols = (size_t)a.ncols;
    // MASKED: the case bytes of this step's node for tile columns `lane` and `lane + 32` (sorted column space)
    const uint8_t *case_a = MASKED ? a.caseb + tile.col0 + (lane < tile.ncols ? lane : tile.ncols - 1) : nullptr;
    const uint8_t *case_b = MASKED ? a.caseb + tile.col0 + (lane + 32 < tile.ncols ? lane + 32 : tile.ncols - 1) : nullptr;
    auto fetch_case = [&](int i, int half) -> uint8_t {
        const int po = XPROG + (i & (XPB * 32 - 1)) * (int)sizeof(Step);
        const int ent = XSTEP(po, ent);
        return *((half == 0 ? case_a : case_b) + (size_t)(unsigned)ent * ncols);
    };
    // lane l fetches the observed states of tile columns l and l + 32
    const uint8_t *obs_a = a.obs + a.orig_col[tile.col0 + (lane < tile.ncols ? lane : tile.ncols - 1)];
    const uint8_t *obs_b = a.obs + a.orig_col[tile.col0 + (lane + 32 < tile.ncols ? lane + 32 : tile.ncols - 1)];
    int chunk_ready = -1;
    auto need_chunk = [&](int j) {
        while (chunk_ready < j) {
            ++chunk_ready;
            mbar_wait(g.bar_prog + (chunk_ready & (XPB - 1)) * 8, (chunk_ready >> 2) & 1);
        }
    };
    // observed states of childless child e of step i for tile columns `lane` (half 0) and `lane + 32` (half 1).
    // The loaded byte must not be touched before it is stored XPD steps later: any arithmetic on it here would
    // make this warp wait for DRAM once per step (it did: profiles/r1_v18 -- one step per memory round trip).
    auto fetch_obs = [&](int i, int e, int half) -> uint8_t {
        const int po = XPROG + (i & (XPB * 32 - 1)) * (int)sizeof(Step);
        const int cn = e == 0 ? XSTEP(po, c_node[0]) : XSTEP(po, c_node[1]);
        const int ce = e == 0 ? XSTEP(po, c_ent[0]) : XSTEP(po, c_ent[1]);
        const bool leaf = cn >= 0 && ce < 0;  // childless child
        const uint8_t *src = (half == 0 ? obs_a : obs_b) + (size_t)(unsigned)(leaf ? cn : 0) * ncols;
        return leaf ? *src : (uint8_t)BNK_NONE;
    };
    need_chunk(0);
    uint8_t o[2][2][XPD];  // [child][column half][ring position]
    uint8_t cb[MASKED ? 2 : 1][XPD];
#pragma unroll
    for (int k = 0; k < XPD; k++) {
#pragma unroll
        for (int e = 0; e < 2; e++)
#pragma unroll
            for (int hf = 0; hf < 2; hf++) o[e][hf][k] = k < n_steps ? fetch_obs(k, e, hf) : (uint8_t)BNK_NONE;
        if (MASKED) {
#pragma unroll
            for (int hf = 0; hf < 2; hf++) cb[hf][k] = k < n_steps ? fetch_case(k, hf) : (uint8_t)0;
        }
    }
    int s = 0, ph = 0;
    for (int i0 = 0; i0 < n_steps; i0 += XPD) {
#pragma unroll
        for (int k = 0; k < XPD; k++) {
            const int i = i0 + k;
            if (i < n_steps) {
                if (i >= ns) mbar_wait(g.bar_empty + s * 8, ph ^ 1);
                const int so = XSTAGE0 + s * g.stage_bytes;
                smem_raw[so + lane] = o[0][0][k];
                smem_raw[so + 32 + lane] = o[0][1][k];
                smem_raw[so + 64 + lane] = o[1][0][k];
                smem_raw[so + 96 + lane] = o[1][1][k];
                if (MASKED) {
                    smem_raw[so + 128 + lane] = cb[0][k];
                    smem_raw[so + 160 + lane] = cb[1][k];
                }
                mbar_arrive(g.bar_full + s * 8);
                const int ip = i + XPD;
                if (ip < n_steps) {
                    need_chunk(ip >> 5);
#pragma unroll
                    for (int e = 0; e < 2; e++)
#pragma unroll
                        for (int hf = 0; hf < 2; hf++) o[e][hf][k] = fetch_obs(ip, e, hf);
                    if (MASKED) {
#pragma unroll
                        for (int hf = 0; hf < 2; hf++) cb[hf][k] = fetch_case(ip, hf);
                    }
                }
                if (++s == ns) { s = 0; ph ^= 1; }
            }
        }
    }
}

// Lane geometry: lane = 5 c + q (c = 0..5, q = 0..4; lanes 30, 31 shadow lanes 25, 26).  Lane q owns the
// outputs {2q, 2q+1} and {10+2q, 11+2q} of C columns: tile columns warp * 6C + j * 6 + c, j < C.
template <int C>
struct XLane {
    int c, q, p0, p1;      // p0 = 2q, p1 = 10 + 2q
    int lc[C], col[C], ocol[C];
    __device__ XLane(const WalkArgs &a, const TileDesc &tile, int warp, int lane)
    {
        c = lane < 30 ? lane / 5 : 5;
        q = lane < 30 ? lane - 5 * c : lane - 30;
        p0 = 2 * q; p1 = 10 + 2 * q;
#pragma unroll
        for (int j = 0; j < C; j++) {
            lc[j] = (warp * C + j) * XCW + c;
            if (lc[j] >= tile.ncols) lc[j] = tile.ncols - 1;  // spare lanes shadow the tile's last column
            col[j] = tile.col0 + lc[j];
            ocol[j] = a.orig_col[col[j]];
        }
    }
};

// a lane's four entries of a 20-vector stored at byte offset `off`: v[0..1] = entries p0.., v[2..3] = entries p1..
#define XLD4(v, off, q)                                                            \
    do {                                                                           \
        const double2 u_ = XS_D2((off) + (q) * 16), w_ = XS_D2((off) + 80 + (q) * 16); \
        (v)[0] = u_.x; (v)[1] = u_.y; (v)[2] = w_.x; (v)[3] = w_.y;                \
    } while (0)
#define XLD4P(v, off, q, plane)                                                        \
    do {                                                                               \
        const double2 u_ = XS_D2((off) + (q) * 16), w_ = XS_D2((off) + (plane) + (q) * 16); \
        (v)[0] = u_.x; (v)[1] = u_.y; (v)[2] = w_.x; (v)[3] = w_.y;                    \
    } while (0)
#define XST4P(off, q, v, plane)                                           \
    do {                                                                  \
        XS_D2((off) + (q) * 16) = make_double2((v)[0], (v)[1]);           \
        XS_D2((off) + (plane) + (q) * 16) = make_double2((v)[2], (v)[3]); \
    } while (0)
// the pair (entries 2 x2, 2 x2 + 1) of a planar vector
#define XPAIR(off, x2, plane) XS_D2((off) + ((x2) < 5 ? (x2) * 16 : (plane) + ((x2) - 5) * 16))
#define XST4(off, q, v)                                          \
    do {                                                         \
        XS_D2((off) + (q) * 16) = make_double2((v)[0], (v)[1]);  \
        XS_D2((off) + 80 + (q) * 16) = make_double2((v)[2], (v)[3]); \
    } while (0)
}  // namespace

// ------------------------------------------------------------------------------------
// up-sweeps.  MODE 0: max-product in log space (joint), MODE 1: sum-product, linear, rescaled.
// ------------------------------------------------------------------------------------
// MASKED (C = 1): per-column forests.  A pre-pass (walk_case_kernel) classified every (step, column) from the
// presence mask exactly as the general walker does -- KIND_NONE (absent / no BN variable), KIND_ROOT (parent absent:
// the prior joins, the component closes here), KIND_MSG -- and recorded which children are present; the byte rides
// in the stage.  The inner loop stays branch-free: an absent child is the unit vector (a select), a component root
// contracts with the UNIT TABLE kept in shared memory instead of the node's table (a per-lane base address), a
// node without a variable computes harmlessly and nobody reads what it leaves.
template <int MODE, int C, bool MASKED>
__global__ void __launch_bounds__((XMAXW + 2) * 32, 1) up_warp_kernel(const WalkArgs a, const int W, const int ns, const int chb)
{
    static_assert(!MASKED || C == 1, "the masked walker keeps one column per lane");
    constexpr int XOBS = MASKED ? XOBS_M : bnk::XOBS;
    constexpr int XSTAGE0 = bnk::XSTAGE0 + (MASKED ? XTAB : 0);
    constexpr int XUNIT = bnk::XSTAGE0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TileDesc tile = a.tiles[blockIdx.x];
    if (MASKED)
        for (int i = threadIdx.x; i < XK * XK; i += blockDim.x) *reinterpret_cast<double *>(smem_raw + XUNIT + i * 8) = MODE == 0 ? 0.0 : 1.0;
    const XGeom g = x_setup(sbase, W, ns, chb, MASKED);
    if (warp >= W) {
        if (warp == W) x_producer_tma<false, MASKED>(a, tile, g, smem_raw, sbase, ns, lane);
        else x_producer_obs<false, MASKED>(a, tile, g, smem_raw, ns, lane);
        return;
    }
    const XLane<C> L(a, tile, warp, lane);
    const int q = L.q;
    const int n_steps = a.n_steps;
    const size_t ncols = (size_t)a.ncols;
    const int wbase = XSTAGE0 + ns * g.stage_bytes + warp * x_warp_bytes(C, a.smem_slots);
    constexpr int PLANE = XCW * C * 80;  // bytes between the two halves of a planar vector
    const int slot_bytes = 2 * PLANE;
    const int stack0 = wbase + 2 * slot_bytes;  // after the two exchange buffers
    const int dummy = stack0 + a.smem_slots * slot_bytes;
    int own[C];  // this lane's columns inside a stack slot / an exchange buffer
#pragma unroll
    for (int j = 0; j < C; j++) own[j] = (j * XCW + L.c) * 80;
    const double ident = MODE == 0 ? 0.0 : 1.0;
    const double prior4[4] = {a.prior[L.p0], a.prior[L.p0 + 1], a.prior[L.p1], a.prior[L.p1 + 1]};
    int esum[C];
    double logacc[C];
#pragma unroll
    for (int j = 0; j < C; j++) { esum[j] = 0; logacc[j] = 0.0; }

    int s = 0, ph = 0;
    unsigned ready = 0;  // result of the early probe of this step's stage
    for (int i = 0; i < n_steps; i++) {
        if ((i & 31) == 0) mbar_wait(g.bar_prog + ((i >> 5) & (XPB - 1)) * 8, (i >> 7) & 1);
        const int po = XPROG + (i & (XPB * 32 - 1)) * (int)sizeof(Step);
        const int ent = XSTEP(po, ent), slot = XSTEP(po, slot);
        const bool rootstep = XSTEP(po, parent) < 0;
        const int so = XSTAGE0 + s * g.stage_bytes;
        if (!ready) mbar_wait(g.bar_full + s * 8, ph);
        {   // probe the next stage now: the probe's round trip (a few hundred cycles) hides behind this step
            const int sn = s + 1 == ns ? 0 : s + 1;
            ready = i + 1 < n_steps ? mbar_test(g.bar_full + sn * 8, s + 1 == ns ? ph ^ 1 : ph) : 0u;
        }
        // ---- this column's case (MASKED) ----
        int kindc[C];
        bool cpres[C][2];
#pragma unroll
        for (int j = 0; j < C; j++) {
            kindc[j] = rootstep ? KIND_ROOT : KIND_MSG; cpres[j][0] = cpres[j][1] = true;
            if (MASKED) {
                const int cb = smem_raw[so + 128 + L.lc[j]];
                kindc[j] = cb & 3; cpres[j][0] = (cb & 4) != 0; cpres[j][1] = (cb & 8) != 0;
            }
        }
        // ---- the (at most two) children ----
        double cv[C][2][4];
        bool isob[C][2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int cn = e == 0 ? XSTEP(po, c_node[0]) : XSTEP(po, c_node[1]);
            const int cs = e == 0 ? XSTEP(po, c_slot[0]) : XSTEP(po, c_slot[1]);
            const int ce = e == 0 ? XSTEP(po, c_ent[0]) : XSTEP(po, c_ent[1]);
#pragma unroll
            for (int j = 0; j < C; j++) {
                isob[j][e] = false;
                if (MASKED && !cpres[j][e]) {  // absent in this column: the unit vector
#pragma unroll
                    for (int k = 0; k < 4; k++) cv[j][e][k] = ident;
                } else if (cs >= 0) {  // message left on this warp's stack
                    XLD4P(cv[j][e], stack0 + cs * slot_bytes + own[j], q, PLANE);
                } else if (cs == SLOT_WIDE) {  // message a wide level kernel left in HBM, [ent][K][ncols]
                    const double *src = a.msg + (size_t)(unsigned)ce * (ncols * XK) + L.col[j];
                    cv[j][e][0] = src[(size_t)L.p0 * ncols]; cv[j][e][1] = src[(size_t)(L.p0 + 1) * ncols];
                    cv[j][e][2] = src[(size_t)L.p1 * ncols]; cv[j][e][3] = src[(size_t)(L.p1 + 1) * ncols];
                } else if (cn >= 0) {  // childless child: a row of its transposed table, picked by the observed state
                    const int tb = so + XOBS + XTAB + e * g.chb;
                    const int ob = smem_raw[so + e * 64 + L.lc[j]];
                    if (ob != BNK_NONE) {
                        XLD4(cv[j][e], tb + ob * (XK * 8), q);
                        isob[j][e] = true;
                    } else {  // uninstantiated childless node: maximise / marginalise its own state
                        XLD4(cv[j][e], tb, q);
                        for (int x = 1; x < XK; x++) {
                            double t4[4];
                            XLD4(t4, tb + x * (XK * 8), q);
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                if (MODE == 0) cv[j][e][k] = t4[k] > cv[j][e][k] ? t4[k] : cv[j][e][k];
                                else cv[j][e][k] += t4[k];
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++) cv[j][e][k] = ident;
                }
            }
        }
        const int gb = wbase + (i & 1) * slot_bytes;
#pragma unroll
        for (int j = 0; j < C; j++) {
            // a root combines [prior, observed children, unobserved children] (the order the factors sit in the
            // reference's bucket); with one child of each kind that fixes the association
            const bool rootc = MASKED ? kindc[j] == KIND_ROOT : rootstep;
            const bool swap = MODE == 0 && rootc && !isob[j][0] && isob[j][1] && cpres[j][0];
            double G[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double c0 = swap ? cv[j][1][k] : cv[j][0][k], c1 = swap ? cv[j][0][k] : cv[j][1][k];
                if (MODE == 0) G[k] = ((rootc ? prior4[k] : 0.0) + c0) + c1;
                else G[k] = ((rootc ? prior4[k] : 1.0) * c0) * c1;
            }
            XST4P(gb + own[j], q, G, PLANE);
        }
        __syncwarp();
        // ---- contract with the node's table: out[p] = op_x T[x][p] (.) G[x] for the lane's four p ----
        // MASKED: a component root contracts with the unit table (every output = max / sum over G)
        const int tv = (MASKED && kindc[0] == KIND_ROOT) ? XUNIT : so + XOBS;
        const int sdst = slot >= 0 ? stack0 + slot * slot_bytes : dummy;
        if (MODE == 0) {
            double best[C][4];
            int bi[C][4];
#pragma unroll
            for (int j = 0; j < C; j++)
#pragma unroll
                for (int k = 0; k < 4; k++) { best[j][k] = -CUDART_INF; bi[j][k] = 0; }
            if (MASKED || !rootstep) {
#pragma unroll
                for (int x2 = 0; x2 < XK / 2; x2++) {
                    double2 gg[C];
#pragma unroll
                    for (int j = 0; j < C; j++) gg[j] = XPAIR(gb + own[j], x2, PLANE);
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int x = 2 * x2 + h;
                        double t4[4];
                        XLD4(t4, tv + x * (XK * 8), q);
#pragma unroll
                        for (int j = 0; j < C; j++) {
                            const double gx = h == 0 ? gg[j].x : gg[j].y;
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                const double val = t4[k] + gx;
                                const bool t = val > best[j][k];
                                best[j][k] = t ? val : best[j][k];
                                bi[j][k] = t ? x : bi[j][k];
                            }
                        }
                    }
                }
            } else {  // unit table: every output is max_x G[x]
#pragma unroll
                for (int j = 0; j < C; j++) {
                    double b = -CUDART_INF;
                    int bx = 0;
#pragma unroll
                    for (int x2 = 0; x2 < XK / 2; x2++) {
                        const double2 gg = XPAIR(gb + own[j], x2, PLANE);
                        const double v0 = 0.0 + gg.x, v1 = 0.0 + gg.y;
                        if (v0 > b) { b = v0; bx = 2 * x2; }
                        if (v1 > b) { b = v1; bx = 2 * x2 + 1; }
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) { best[j][k] = b; bi[j][k] = bx; }
                }
            }
#pragma unroll
            for (int j = 0; j < C; j++) {
                XST4P(sdst + own[j], q, best[j], PLANE);
                uint8_t *pr = a.ptr + ((size_t)(unsigned)ent * ncols + L.col[j]) * XK;
                *reinterpret_cast<uint16_t *>(pr + L.p0) = (uint16_t)(bi[j][0] | (bi[j][1] << 8));
                *reinterpret_cast<uint16_t *>(pr + L.p1) = (uint16_t)(bi[j][2] | (bi[j][3] << 8));
            }
        } else {
            double acc[C][4], acb[C][4];
            int mhi[C];
#pragma unroll
            for (int j = 0; j < C; j++) {
                mhi[j] = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) { acc[j][k] = 0.0; acb[j][k] = 0.0; }
            }
            if (MASKED || !rootstep) {
#pragma unroll
                for (int x2 = 0; x2 < XK / 2; x2++) {
                    double ta[4], tb4[4];
                    XLD4(ta, tv + (2 * x2) * (XK * 8), q);
                    XLD4(tb4, tv + (2 * x2 + 1) * (XK * 8), q);
#pragma unroll
                    for (int j = 0; j < C; j++) {
                        const double2 gg = XPAIR(gb + own[j], x2, PLANE);
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            acc[j][k] = fma(ta[k], gg.x, acc[j][k]);
                            acb[j][k] = fma(tb4[k], gg.y, acb[j][k]);
                        }
                        mhi[j] = max(mhi[j], max(__double2hiint(gg.x), __double2hiint(gg.y)));
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < C; j++) {
#pragma unroll
                    for (int x2 = 0; x2 < XK / 2; x2++) {
                        const double2 gg = XPAIR(gb + own[j], x2, PLANE);
                        acc[j][0] += gg.x; acb[j][0] += gg.y;
                        mhi[j] = max(mhi[j], max(__double2hiint(gg.x), __double2hiint(gg.y)));
                    }
#pragma unroll
                    for (int k = 1; k < 4; k++) { acc[j][k] = acc[j][0]; acb[j][k] = acb[j][0]; }
                }
            }
#pragma unroll
            for (int j = 0; j < C; j++) {
                int e;
                const double scale = x_pow2_scale(mhi[j], e);
                double m[4];
#pragma unroll
                for (int k = 0; k < 4; k++) m[k] = (acc[j][k] + acb[j][k]) * scale;
                esum[j] += (MASKED && kindc[j] == KIND_NONE) ? 0 : e;
                XST4P(sdst + own[j], q, m, PLANE);
                if (MASKED ? kindc[j] == KIND_ROOT : rootstep) logacc[j] += log(m[0]);
                else if (a.msg && (!MASKED || kindc[j] == KIND_MSG)) {
                    double *dst = a.msg + ((size_t)(unsigned)ent * ncols + L.col[j]) * XK;
                    *reinterpret_cast<double2 *>(dst + L.p0) = make_double2(m[0], m[1]);
                    *reinterpret_cast<double2 *>(dst + L.p1) = make_double2(m[2], m[3]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(g.bar_empty + s * 8);
        if (++s == ns) { s = 0; ph ^= 1; }
    }
    if (MODE == 1 && a.out_logl && q == 0) {
#pragma unroll
        for (int j = 0; j < C; j++) {
            const int ew = a.esum_wide ? a.esum_wide[L.col[j]] : 0;
            const double lw = (MASKED && a.logl_wide) ? a.logl_wide[L.col[j]] : 0.0;  // components closed inside the wide levels
            a.out_logl[L.ocol[j]] = logacc[j] + lw + (double)(esum[j] + ew) * 0.693147180559945309417232121458;
        }
    }
}

// ------------------------------------------------------------------------------------
// down-sweep (see sweeps.cu for the recurrences): from-above vector D = P^T T, posterior
// D * prod(children) normalised, T_child = D * prod(siblings)
// ------------------------------------------------------------------------------------
template <int C, bool MASKED>
__global__ void __launch_bounds__((XMAXW + 2) * 32, 1) down_warp_kernel(const WalkArgs a, const int W, const int ns, const int chb)
{
    static_assert(!MASKED || C == 1, "the masked walker keeps one column per lane");
    constexpr int XOBS = MASKED ? XOBS_M : bnk::XOBS;
    constexpr int XSTAGE0 = bnk::XSTAGE0 + (MASKED ? XTAB : 0);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TileDesc tile = a.tiles[blockIdx.x];
    const XGeom g = x_setup(sbase, W, ns, chb, MASKED);
    if (warp >= W) {
        if (warp == W) x_producer_tma<true, MASKED>(a, tile, g, smem_raw, sbase, ns, lane);
        else x_producer_obs<true, MASKED>(a, tile, g, smem_raw, ns, lane);
        return;
    }
    const XLane<C> L(a, tile, warp, lane);
    const int q = L.q;
    const int n_steps = a.n_steps;
    const size_t ncols = (size_t)a.ncols;
    const int wbase = XSTAGE0 + ns * g.stage_bytes + warp * x_warp_bytes(C, a.smem_slots);
    constexpr int PLANE = XCW * C * 80;
    const int slot_bytes = 2 * PLANE;
    const int stack0 = wbase + 2 * slot_bytes;
    const int dummy = stack0 + a.smem_slots * slot_bytes;
    int own[C];
#pragma unroll
    for (int j = 0; j < C; j++) own[j] = (j * XCW + L.c) * 80;
    const double prior4[4] = {a.prior[L.p0], a.prior[L.p0 + 1], a.prior[L.p1], a.prior[L.p1 + 1]};

    int s = 0, ph = 0;
    unsigned ready = 0;  // result of the early probe of this step's stage
    for (int i = 0; i < n_steps; i++) {
        if ((i & 31) == 0) mbar_wait(g.bar_prog + ((i >> 5) & (XPB - 1)) * 8, (i >> 7) & 1);
        const int po = XPROG + (i & (XPB * 32 - 1)) * (int)sizeof(Step);
        const int ent = XSTEP(po, ent), slot = XSTEP(po, slot);
        const bool rootstep = XSTEP(po, parent) < 0;
        const int so = XSTAGE0 + s * g.stage_bytes;
        const int tsrc = slot >= 0 ? stack0 + slot * slot_bytes : dummy;
        if (!ready) mbar_wait(g.bar_full + s * 8, ph);
        {   // probe the next stage now: the probe's round trip (a few hundred cycles) hides behind this step
            const int sn = s + 1 == ns ? 0 : s + 1;
            ready = i + 1 < n_steps ? mbar_test(g.bar_full + sn * 8, s + 1 == ns ? ph ^ 1 : ph) : 0u;
        }
        // ---- this column's case (MASKED): the classification of the up-sweep ----
        int kindc[C];
        bool cpres[C][2];
#pragma unroll
        for (int j = 0; j < C; j++) {
            kindc[j] = rootstep ? KIND_ROOT : KIND_MSG; cpres[j][0] = cpres[j][1] = true;
            if (MASKED) {
                const int cb = smem_raw[so + 128 + L.lc[j]];
                kindc[j] = cb & 3; cpres[j][0] = (cb & 4) != 0; cpres[j][1] = (cb & 8) != 0;
            }
        }
        // ---- from-above vector of this node: D[x] = sum_p P[p][x] T[p], scaled by a power of two ----
        double D[C][4];
        if (MASKED || !rootstep) {
            const int tv = so + XOBS;
            double acc[C][4], acb[C][4];
            int mhi[C];
#pragma unroll
            for (int j = 0; j < C; j++) {
                mhi[j] = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) { acc[j][k] = 0.0; acb[j][k] = 0.0; }
            }
#pragma unroll
            for (int x2 = 0; x2 < XK / 2; x2++) {
                double ta[4], tb4[4];
                XLD4(ta, tv + (2 * x2) * (XK * 8), q);
                XLD4(tb4, tv + (2 * x2 + 1) * (XK * 8), q);
#pragma unroll
                for (int j = 0; j < C; j++) {
                    const double2 tt = XPAIR(tsrc + own[j], x2, PLANE);
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        acc[j][k] = fma(ta[k], tt.x, acc[j][k]);
                        acb[j][k] = fma(tb4[k], tt.y, acb[j][k]);
                    }
                    mhi[j] = max(mhi[j], max(__double2hiint(tt.x), __double2hiint(tt.y)));
                }
            }
#pragma unroll
            for (int j = 0; j < C; j++) {
                int e;
                const double scale = x_pow2_scale(mhi[j], e);
#pragma unroll
                for (int k = 0; k < 4; k++) D[j][k] = (acc[j][k] + acb[j][k]) * scale;
                if (MASKED && kindc[j] != KIND_MSG) {  // a component root starts from the prior (what the table gave is discarded)
#pragma unroll
                    for (int k = 0; k < 4; k++) D[j][k] = prior4[k];
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < C; j++)
#pragma unroll
                for (int k = 0; k < 4; k++) D[j][k] = prior4[k];
        }
        // ---- children: stored up-message (internal), table row (observed childless), row sums, or 1 ----
        double cv[C][2][4];
        int cslot[2], cent[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int cn = e == 0 ? XSTEP(po, c_node[0]) : XSTEP(po, c_node[1]);
            const int cs = e == 0 ? XSTEP(po, c_slot[0]) : XSTEP(po, c_slot[1]);
            const int ce = e == 0 ? XSTEP(po, c_ent[0]) : XSTEP(po, c_ent[1]);
            cslot[e] = cs; cent[e] = ce;
            const int area = so + XOBS + XTAB + e * g.chb;
#pragma unroll
            for (int j = 0; j < C; j++) {
                if (MASKED && !cpres[j][e]) {  // absent in this column: the unit vector
                    cv[j][e][0] = cv[j][e][1] = cv[j][e][2] = cv[j][e][3] = 1.0;
                } else if (cs == SLOT_WIDE) {
                    const double *src = a.msg + (size_t)(unsigned)ce * (ncols * XK) + L.col[j];
                    cv[j][e][0] = src[(size_t)L.p0 * ncols]; cv[j][e][1] = src[(size_t)(L.p0 + 1) * ncols];
                    cv[j][e][2] = src[(size_t)L.p1 * ncols]; cv[j][e][3] = src[(size_t)(L.p1 + 1) * ncols];
                } else if (ce >= 0) {  // walker-internal child: its up-message came with the stage, [tile column][K]
                    XLD4(cv[j][e], area + L.lc[j] * (XK * 8), q);
                } else if (cn >= 0) {
                    const int ob = smem_raw[so + e * 64 + L.lc[j]];
                    if (ob != BNK_NONE) {
                        XLD4(cv[j][e], area + ob * (XK * 8), q);
                    } else {
                        cv[j][e][0] = cv[j][e][1] = cv[j][e][2] = cv[j][e][3] = 0.0;
                        for (int x = 0; x < XK; x++) {
                            double t4[4];
                            XLD4(t4, area + x * (XK * 8), q);
#pragma unroll
                            for (int k = 0; k < 4; k++) cv[j][e][k] += t4[k];
                        }
                    }
                } else {
                    cv[j][e][0] = cv[j][e][1] = cv[j][e][2] = cv[j][e][3] = 1.0;
                }
            }
        }
        const int gb = wbase + (i & 1) * slot_bytes;
        double U[C][4], T0[C][4], T1[C][4];
#pragma unroll
        for (int j = 0; j < C; j++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                U[j][k] = D[j][k] * (cv[j][0][k] * cv[j][1][k]);
                T0[j][k] = D[j][k] * cv[j][1][k];
                T1[j][k] = D[j][k] * cv[j][0][k];
                if (MASKED && kindc[j] == KIND_NONE) U[j][k] = CUDART_NAN;  // no BN variable: no posterior
            }
            XST4P(gb + own[j], q, U[j], PLANE);
        }
        // every lane has read this node's from-above vector before the barrier; the children's slots never alias
        // this step's own slot (tree_plan.cpp), and they are written after it
        __syncwarp();
        {
            const int d0 = cslot[0] >= 0 ? stack0 + cslot[0] * slot_bytes : dummy;
            const int d1 = cslot[1] >= 0 ? stack0 + cslot[1] * slot_bytes : dummy;
#pragma unroll
            for (int j = 0; j < C; j++) {
                XST4P(d0 + own[j], q, T0[j], PLANE);
                XST4P(d1 + own[j], q, T1[j], PLANE);
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    if (cslot[e] == SLOT_WIDE && (!MASKED || (cpres[j][e] && kindc[j] != KIND_NONE))) {  // handed to the wide level kernels through HBM, [ent][K][ncols]
                        double *dst = a.tmsg + (size_t)(unsigned)cent[e] * (ncols * XK) + L.col[j];
                        const double *Tv = e == 0 ? T0[j] : T1[j];
                        dst[(size_t)L.p0 * ncols] = Tv[0]; dst[(size_t)(L.p0 + 1) * ncols] = Tv[1];
                        dst[(size_t)L.p1 * ncols] = Tv[2]; dst[(size_t)(L.p1 + 1) * ncols] = Tv[3];
                    }
                }
            }
        }
        // ---- normalise and emit the posterior ----
        const int q0 = a.qrow ? a.qrow[ent] : ent;
        if (a.out_post && q0 >= 0) {
#pragma unroll
            for (int j = 0; j < C; j++) {
                double sum = 0.0;
#pragma unroll
                for (int x2 = 0; x2 < XK / 2; x2++) { const double2 gg = XPAIR(gb + own[j], x2, PLANE); sum += gg.x; sum += gg.y; }
                const double r = x_rcp(sum);
                double *dst = a.out_post + ((size_t)(unsigned)q0 * ncols + L.ocol[j]) * XK;
                *reinterpret_cast<double2 *>(dst + L.p0) = make_double2(U[j][0] * r, U[j][1] * r);
                *reinterpret_cast<double2 *>(dst + L.p1) = make_double2(U[j][2] * r, U[j][3] * r);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(g.bar_empty + s * 8);
        if (++s == ns) { s = 0; ph ^= 1; }
    }
}

// ------------------------------------------------------------------------------------
// Sum-product sweeps on the FP64 tensor-core path (DMMA.8x8x4, mma.sync m8n8k4).
//
// The contraction of a step, out[p][col] = sum_x T[p][x] * G[x][col], is a 20 x 20 by 20 x 8 matrix product once a
// warp owns 8 columns.  What that buys here is not flops but shared-memory traffic, the walkers' bound: an A
// fragment is ONE 8-byte load per lane (conflict-free on the raw table layout) and feeds all 8 columns, and the
// B fragments -- G itself -- are built in registers straight from the children's data, so the exchange buffer and
// its 10 LDS.128 per column disappear: ~15 shared-memory wavefronts per column and step instead of ~37.
// Fragment layout (PTX ISA, mma.m8n8k4 .f64): A[row = lane / 4][k = lane % 4], B[k = lane % 4][col = lane / 4],
// C/D[row = lane / 4][col = 2 (lane % 4) + {0, 1}].  Rows 20..23 of the third row tile are padding: they read
// whatever follows the table row and are never stored.  Max-product has no such form; the joint sweep keeps
// up_warp_kernel.  Tolerance-checked (1e-9) like every sum-product kernel; the accumulation order inside a DMMA
// differs from the scalar kernels'.
// ------------------------------------------------------------------------------------
namespace {
constexpr int MCW = 8;  // columns per warp

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#define XS_D(off) (*reinterpret_cast<double *>(smem_raw + (off)))

struct MLane {
    int xr, g;            // lane % 4, lane / 4
    int lcb, colb, ocolb; // the column this lane feeds as B operand (g)
    int lcc[2], colc[2], ocolc[2];  // the two columns it holds results for (2 xr + i)
    __device__ MLane(const WalkArgs &a, const TileDesc &tile, int warp, int lane)
    {
        xr = lane & 3; g = lane >> 2;
        lcb = warp * MCW + g;
        if (lcb >= tile.ncols) lcb = tile.ncols - 1;
        colb = tile.col0 + lcb; ocolb = a.orig_col[colb];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            lcc[i] = warp * MCW + 2 * xr + i;
            if (lcc[i] >= tile.ncols) lcc[i] = tile.ncols - 1;
            colc[i] = tile.col0 + lcc[i]; ocolc[i] = a.orig_col[colc[i]];
        }
    }
};
__host__ __device__ inline int m_warp_bytes(int slots) { return (slots + 1) * MCW * XK * 8; }
}  // namespace

__global__ void __launch_bounds__((XMAXW + 2) * 32, 1) up_mma_kernel(const WalkArgs a, const int W, const int ns, const int chb)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TileDesc tile = a.tiles[blockIdx.x];
    const XGeom g = x_setup(sbase, W, ns, chb);
    if (warp >= W) {
        if (warp == W) x_producer_tma<false>(a, tile, g, smem_raw, sbase, ns, lane);
        else x_producer_obs<false>(a, tile, g, smem_raw, ns, lane);
        return;
    }
    const MLane L(a, tile, warp, lane);
    const int n_steps = a.n_steps;
    const size_t ncols = (size_t)a.ncols;
    const int slot_bytes = MCW * XK * 8;  // [column][20]
    const int stack0 = XSTAGE0 + ns * g.stage_bytes + warp * m_warp_bytes(a.smem_slots);
    const int dummy = stack0 + a.smem_slots * slot_bytes;
    const unsigned FULL = 0xffffffffu;
    double prior5[5];
#pragma unroll
    for (int kt = 0; kt < 5; kt++) prior5[kt] = a.prior[4 * kt + L.xr];
    int esum = 0;        // of column L.colb (identical in the 4 lanes of a group)
    double logacc = 0.0;

    int s = 0, ph = 0;
    unsigned ready = 0;  // result of the early probe of this step's stage
    for (int i = 0; i < n_steps; i++) {
        if ((i & 31) == 0) mbar_wait(g.bar_prog + ((i >> 5) & (XPB - 1)) * 8, (i >> 7) & 1);
        const int po = XPROG + (i & (XPB * 32 - 1)) * (int)sizeof(Step);
        const int ent = XSTEP(po, ent), slot = XSTEP(po, slot);
        const bool rootstep = XSTEP(po, parent) < 0;
        const int so = XSTAGE0 + s * g.stage_bytes;
        if (!ready) mbar_wait(g.bar_full + s * 8, ph);
        {   // probe the next stage now: the probe's round trip (a few hundred cycles) hides behind this step
            const int sn = s + 1 == ns ? 0 : s + 1;
            ready = i + 1 < n_steps ? mbar_test(g.bar_full + sn * 8, s + 1 == ns ? ph ^ 1 : ph) : 0u;
        }
        // ---- G as B fragments: entries x = 4 kt + xr of column g ----
        double G[5];
#pragma unroll
        for (int kt = 0; kt < 5; kt++) G[kt] = rootstep ? prior5[kt] : 1.0;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int cn = e == 0 ? XSTEP(po, c_node[0]) : XSTEP(po, c_node[1]);
            const int cs = e == 0 ? XSTEP(po, c_slot[0]) : XSTEP(po, c_slot[1]);
            const int ce = e == 0 ? XSTEP(po, c_ent[0]) : XSTEP(po, c_ent[1]);
            double cv[5];
            if (cs >= 0) {
                const int src = stack0 + cs * slot_bytes + L.g * (XK * 8) + L.xr * 8;
#pragma unroll
                for (int kt = 0; kt < 5; kt++) cv[kt] = XS_D(src + kt * 32);
            } else if (cs == SLOT_WIDE) {
                const double *src = a.msg + (size_t)(unsigned)ce * (ncols * XK) + L.colb;
#pragma unroll
                for (int kt = 0; kt < 5; kt++) cv[kt] = src[(size_t)(4 * kt + L.xr) * ncols];
            } else if (cn >= 0) {
                const int tb = so + XOBS + XTAB + e * g.chb + L.xr * 8;
                const int ob = smem_raw[so + e * 64 + L.lcb];
                if (ob != BNK_NONE) {
#pragma unroll
                    for (int kt = 0; kt < 5; kt++) cv[kt] = XS_D(tb + ob * (XK * 8) + kt * 32);
                } else {  // uninstantiated childless node: marginalise its own state
#pragma unroll
                    for (int kt = 0; kt < 5; kt++) cv[kt] = 0.0;
                    for (int x = 0; x < XK; x++) {
#pragma unroll
                        for (int kt = 0; kt < 5; kt++) cv[kt] += XS_D(tb + x * (XK * 8) + kt * 32);
                    }
                }
            } else {
#pragma unroll
                for (int kt = 0; kt < 5; kt++) cv[kt] = 1.0;
            }
#pragma unroll
            for (int kt = 0; kt < 5; kt++) G[kt] *= cv[kt];
        }
        int mhi = 0;
#pragma unroll
        for (int kt = 0; kt < 5; kt++) mhi = max(mhi, __double2hiint(G[kt]));
        mhi = max(mhi, __shfl_xor_sync(FULL, mhi, 1));
        mhi = max(mhi, __shfl_xor_sync(FULL, mhi, 2));
        int eb;
        const double scale_b = x_pow2_scale(mhi, eb);  // of column g
        esum += eb;
        if (!rootstep) {
            // ---- out[p][col] = sum_x T[x][p] G[x][col]: 3 row tiles x 5 k tiles ----
            const int tv = so + XOBS + L.xr * (XK * 8) + L.g * 8;
            double D[3][2];
#pragma unroll
            for (int mt = 0; mt < 3; mt++) {  // two accumulator chains per row tile (k tiles 0, 2, 4 and 1, 3): shorter dependent chains
                double E0 = 0.0, E1 = 0.0;
                D[mt][0] = 0.0; D[mt][1] = 0.0;
#pragma unroll
                for (int kt = 0; kt < 5; kt++) {
                    const double av = XS_D(tv + kt * (4 * XK * 8) + mt * 64);
                    if (kt & 1) dmma(E0, E1, av, G[kt]);
                    else dmma(D[mt][0], D[mt][1], av, G[kt]);
                }
                D[mt][0] += E0; D[mt][1] += E1;
            }
            double sc[2];
            sc[0] = __shfl_sync(FULL, scale_b, 8 * L.xr);      // lane 4 * (2 xr)
            sc[1] = __shfl_sync(FULL, scale_b, 8 * L.xr + 4);  // lane 4 * (2 xr + 1)
            const int sdst = slot >= 0 ? stack0 + slot * slot_bytes : dummy;
#pragma unroll
            for (int mt = 0; mt < 3; mt++) {
                const int p = 8 * mt + L.g;
                if (p < XK) {
#pragma unroll
                    for (int c2 = 0; c2 < 2; c2++) {
                        const double m = D[mt][c2] * sc[c2];
                        XS_D(sdst + (2 * L.xr + c2) * (XK * 8) + p * 8) = m;
                        if (a.msg) a.msg[((size_t)(unsigned)ent * ncols + L.colc[c2]) * XK + p] = m;
                    }
                }
            }
        } else {  // unit table: the column's total, into the log-likelihood
            double tot = 0.0;
#pragma unroll
            for (int kt = 0; kt < 5; kt++) tot += G[kt];
            tot += __shfl_xor_sync(FULL, tot, 1);
            tot += __shfl_xor_sync(FULL, tot, 2);
            logacc += log(tot * scale_b);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(g.bar_empty + s * 8);
        if (++s == ns) { s = 0; ph ^= 1; }
    }
    if (a.out_logl && L.xr == 0) {
        const int ew = a.esum_wide ? a.esum_wide[L.colb] : 0;
        a.out_logl[L.ocolb] = logacc + (double)(esum + ew) * 0.693147180559945309417232121458;
    }
}

// ------------------------------------------------------------------------------------
int warp_walk_cols_per_warp(int C) { return XCW * C; }
size_t warp_walk_smem_bytes(int W, int C, int ns, int slots, bool down, bool masked)
{
    return (size_t)x_stage0(masked) + (size_t)ns * x_stage_bytes(x_child_bytes(W * XCW * C, down), masked) + (size_t)W * x_warp_bytes(C, slots);
}

template <typename KernelT>
static cudaError_t launch_warp(KernelT kern, const WalkArgs &a, int n_tiles, int W, int C, int ns, bool down, cudaStream_t st)
{
    if (W < 1 || W > XMAXW || ns < 2 || ns > XMAXNS || W * XCW * C > 64) return cudaErrorInvalidValue;
    if (a.caseb && C != 1) return cudaErrorInvalidValue;  // the masked walker keeps one column per lane
    const size_t smem = warp_walk_smem_bytes(W, C, ns, a.smem_slots, down, a.caseb != nullptr);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    kern<<<n_tiles, (W + 2) * 32, smem, st>>>(a, W, ns, x_child_bytes(W * XCW * C, down));
    return cudaGetLastError();
}
cudaError_t launch_joint_up_warp(const WalkArgs &a, int n_tiles, int W, int C, int ns, cudaStream_t st)
{
    if (a.caseb) return launch_warp(up_warp_kernel<0, 1, true>, a, n_tiles, W, C, ns, false, st);
    if (C == 1) return launch_warp(up_warp_kernel<0, 1, false>, a, n_tiles, W, C, ns, false, st);
    if (C == 2) return launch_warp(up_warp_kernel<0, 2, false>, a, n_tiles, W, C, ns, false, st);
    if (C == 3) return launch_warp(up_warp_kernel<0, 3, false>, a, n_tiles, W, C, ns, false, st);
    return cudaErrorInvalidValue;
}
cudaError_t launch_sum_up_warp(const WalkArgs &a, int n_tiles, int W, int C, int ns, cudaStream_t st)
{
    if (a.caseb) return launch_warp(up_warp_kernel<1, 1, true>, a, n_tiles, W, C, ns, false, st);
    if (C == 1) return launch_warp(up_warp_kernel<1, 1, false>, a, n_tiles, W, C, ns, false, st);
    if (C == 2) return launch_warp(up_warp_kernel<1, 2, false>, a, n_tiles, W, C, ns, false, st);
    if (C == 3) return launch_warp(up_warp_kernel<1, 3, false>, a, n_tiles, W, C, ns, false, st);
    return cudaErrorInvalidValue;
}
cudaError_t launch_sum_down_warp(const WalkArgs &a, int n_tiles, int W, int C, int ns, cudaStream_t st)
{
    if (a.caseb) return launch_warp(down_warp_kernel<1, true>, a, n_tiles, W, C, ns, true, st);
    if (C == 1) return launch_warp(down_warp_kernel<1, false>, a, n_tiles, W, C, ns, true, st);
    if (C == 2) return launch_warp(down_warp_kernel<2, false>, a, n_tiles, W, C, ns, true, st);
    if (C == 3) return launch_warp(down_warp_kernel<3, false>, a, n_tiles, W, C, ns, true, st);
    return cudaErrorInvalidValue;
}


// Pre-pass of the masked walker: the case byte of every (walk step, sorted column) from the presence mask
// (original column order) -- kind | child 0 present << 2 | child 1 present << 3, where kind follows the general
// walker (sweeps.cu): KIND_NONE when the node is absent, or root-like without a present child (no BN variable:
// PhyloBN.create gives variables to connected nodes only); KIND_ROOT when its parent is absent (or it has none);
// else KIND_MSG.  Row `ent` of out ([n_int][ncols]) is the kind array the traceback reads (it masks with 3).
__global__ void __launch_bounds__(256) walk_case_kernel(const Step *__restrict__ prog, int n_steps, const uint8_t *__restrict__ present,
                                                        const int32_t *__restrict__ orig_col, int ncols, uint8_t *__restrict__ out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    const int oc = orig_col[c];
    for (int i = blockIdx.y; i < n_steps; i += gridDim.y) {
        const Step &sp = prog[i];
        const bool pv = present[(size_t)sp.node * ncols + oc] != 0;
        const bool pp = sp.parent >= 0 && present[(size_t)sp.parent * ncols + oc] != 0;
        const bool c0 = sp.c_node[0] >= 0 && present[(size_t)sp.c_node[0] * ncols + oc] != 0;
        const bool c1 = sp.c_node[1] >= 0 && present[(size_t)sp.c_node[1] * ncols + oc] != 0;
        const int kind = !pv ? KIND_NONE : (pp ? KIND_MSG : ((c0 || c1) ? KIND_ROOT : KIND_NONE));
        out[(size_t)sp.ent * ncols + c] = (uint8_t)(kind | (c0 ? 4 : 0) | (c1 ? 8 : 0));
    }
}
cudaError_t launch_walk_case(const Step *prog, int n_steps, const uint8_t *present, const int32_t *orig_col, int ncols,
                             uint8_t *out, cudaStream_t st)
{
    if (n_steps <= 0 || ncols <= 0) return cudaSuccess;
    dim3 grid((ncols + 255) / 256, n_steps < 2048 ? n_steps : 2048);
    walk_case_kernel<<<grid, 256, 0, st>>>(prog, n_steps, present, orig_col, ncols, out);
    return cudaGetLastError();
}

// DMMA variants: 8 columns per warp, tiles of at most 8 W columns
int mma_walk_cols_per_warp() { return MCW; }
size_t mma_walk_smem_bytes(int W, int ns, int slots, bool down)
{
    return (size_t)XSTAGE0 + (size_t)ns * x_stage_bytes(x_child_bytes(W * MCW, down)) + (size_t)W * m_warp_bytes(slots);
}
template <typename KernelT>
static cudaError_t launch_mma(KernelT kern, const WalkArgs &a, int n_tiles, int W, int ns, bool down, cudaStream_t st)
{
    if (W < 1 || W > XMAXW || ns < 2 || ns > XMAXNS || W * MCW > 64) return cudaErrorInvalidValue;
    const size_t smem = mma_walk_smem_bytes(W, ns, a.smem_slots, down);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    kern<<<n_tiles, (W + 2) * 32, smem, st>>>(a, W, ns, x_child_bytes(W * MCW, down));
    return cudaGetLastError();
}
cudaError_t launch_sum_up_mma(const WalkArgs &a, int n_tiles, int W, int ns, cudaStream_t st) { return launch_mma(up_mma_kernel, a, n_tiles, W, ns, false, st); }

}  // namespace bnk
