// indel.cu -- gap-augmented peeling on the device: asr.IndelPeeler (Rivas & Eddy 2008) for every alignment column,
// and the optimisation loop built on it (the reference's real "C5").
//
//   IndelPeeler.decorate / felsensteinsExtendedPeeling      src/asr/IndelPeeler.java:229-356
//   containsGap                                             :359-390
//   computeColumnPriors (one pass per indel-rate category)  :93-123
//   calcProbAlnGivenTree, probExtraCol                      :125-193
//   optimiseMuLambda + LikelihoodEvaluator (Brent, mu = lambda) :469-560, src/smile/math/special/Minimise.java:8-107
//   bn.ctmc.GapSubstModel                                   src/bn/ctmc/GapSubstModel.java:26-131, :166-247
//
// The reference runs one IndelPeeler per column on a thread pool, in log space: per (column, node, parent residue,
// child) one logsumexp over 20-21 terms, each term a Math.log of a product -- ~17,000 log + 17,000 exp per node
// and column.  Here a THREAD owns a column and a node's vector lives in registers, in LINEAR space with a
// power-of-two exponent per (node, column) for the residue vector and another for the gap entry (the two differ
// by thousands of binades as soon as one child is an all-gap clade and its sibling is not):
//     R_v[p]  = prod_c ( sum_x A_c[p][x] R_c[x]  +  [c all gaps] del_c )
//     Gap_v   = sum_c [c all gaps] ( sum_q R_c[q] ins_v[q] + Gap_c )          (a SUM over children, as the reference has it)
//     column  = logsumexp( log Gap_root,  log(geom) + log sum_r F[r] R_root[r] )
// with the per-branch tables A / del / ins of indel_tables.cu.  Messages travel through HBM as
// [ancestor][state][column] rows (coalesced loads and stores).  Schedule: height levels with >= LEVEL_MIN_NODES
// nodes are one launch each (grid = nodes x column tiles); the narrow rest of the tree -- an upward-closed set --
// is walked by the same column-owning threads in one launch, children before parents (decreasing index).
// Parity: oracle/bnk_indel.c (log space, literal), 1e-9 relative on every column's log-likelihood.
#include "indel.cuh"
#include "workspace.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <thread>

namespace bnk {
namespace {

constexpr int IND_THREADS = 128;       // columns per CTA tile of the level kernel
constexpr int LEVEL_TPC = 8;           // column tiles a level CTA walks with its children's tables staged once
constexpr int WALK_THREADS = 128;      // columns per CTA of the walk kernel (measured: 64 ms on the 10k-level caterpillar; 122 ms with one-warp CTAs)
constexpr int LEVEL_MIN_NODES = 8;     // narrower height levels go to the walk kernel
constexpr int E_ZERO = INT32_MIN / 2;  // exponent marking an exactly-zero gap entry

struct IndelArgs {
    const int32_t *first, *kids, *ent_of_node, *parent;
    const uint8_t *obs;   // [node][ncols], BNK_NONE = gap; childless rows only
    int32_t ncols;
    IndelTables tb;
    const double *F;      // K + 1
    double *msg;          // [ent][K + 1][ncols]: residue mantissas (max in [1, 2)), gap mantissa
    int32_t *eR, *eG;     // [ent][ncols]
    uint8_t *cg;          // [ent][ncols] containsGap: every extant below is a gap
    double log_geom;
    double *out_logl;     // [ncols]
    const int32_t *nodes; // the nodes of this launch (a level) / the narrow nodes in processing order
    int32_t n_launch_nodes;
    const int4 *wdesc;    // walk kernel: two int4 per step {node, ent, n_children, first child index} {child 0, child 1, ent 0, ent 1}
};

__device__ __forceinline__ double pow2i(int e) { return __hiloint2double((1023 + e) << 20, 0); }  // e in [-1022, 1023]
__device__ __forceinline__ int exponent_of(double m) { return ((__double2hiint(m) >> 20) & 0x7ff) - 1023; }

// a * 2^ea + b * 2^eb as (mantissa, exponent); a zero operand carries no exponent
__device__ __forceinline__ void add_scaled(double a, int ea, double b, int eb, double &r, int &er)
{
    if (b == 0.0) { r = a; er = ea; return; }
    if (a == 0.0) { r = b; er = eb; return; }
    if (ea >= eb) {
        const int d = eb - ea;
        r = a + (d < -1000 ? 0.0 : b * pow2i(d));
        er = ea;
    } else {
        const int d = ea - eb;
        r = b + (d < -1000 ? 0.0 : a * pow2i(d));
        er = eb;
    }
}

// Per-thread state of one (node, column) while its children are folded in.
template <int K>
struct PeelState {
    double Rv[K];
    int eRv;
    double gm;
    int ge;
    bool cgv;
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int p = 0; p < K; p++) Rv[p] = 1.0;
        eRv = 0; gm = 0.0; ge = 0; cgv = true;
    }
};

// fold child `ch` (its table A_ch in shared memory at sA, the parent's insertion vector at s_ins) into the state
// prev / prev_ent: the state this thread computed in its previous walk step (registers); a child that IS that node --
// the chain of a deep tree -- is not read back from HBM
// ent_known / ob_known / del_known: values the walk kernel fetched a step ahead (ent_known = INT_MIN: look them up here)
template <int K>
__device__ __forceinline__ void fold_child(const IndelArgs &a, int ch, int col, const double *sA, const double *s_ins, PeelState<K> &st,
                                           const PeelState<K> *prev = nullptr, int prev_ent = -1, int ent_known = INT32_MIN,
                                           uint8_t ob_known = 0, double del_known = 0.0)
{
    const size_t nc = (size_t)a.ncols;
    const bool known = ent_known != INT32_MIN;
    const double del = known ? del_known : a.tb.del[ch];
    const int ent = known ? ent_known : a.ent_of_node[ch];
    if (ent < 0) {   // an extant (calcLeafPeelingProbabilities :273-284)
        const uint8_t o = known ? ob_known : a.obs[(size_t)ch * nc + col];
        if (o == BNK_NONE) {
#pragma unroll
            for (int p = 0; p < K; p++) st.Rv[p] *= del;
        } else {
            st.cgv = false;
#pragma unroll
            for (int p = 0; p < K; p++) st.Rv[p] *= sA[o * K + p];
        }
    } else {
        double Rc[K];
        const bool reuse = prev != nullptr && ent == prev_ent;
        const double *m = a.msg + ((size_t)ent * (K + 1)) * nc + col;
#pragma unroll
        for (int x = 0; x < K; x++) Rc[x] = reuse ? prev->Rv[x] : m[(size_t)x * nc];
        const int eRc = reuse ? prev->eRv : a.eR[(size_t)ent * nc + col];
        const bool cgc = reuse ? prev->cgv : a.cg[(size_t)ent * nc + col] != 0;
        if (cgc) {
            // ancestral gap term of this child (:297-327): sum_q R_c[q] ins_v[q] + Gap_c
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < K; q++) s += Rc[q] * s_ins[q];
            const double Gc = reuse ? prev->gm : m[(size_t)K * nc];
            const int eGc = reuse ? prev->ge : a.eG[(size_t)ent * nc + col];
            double tm;
            int te;
            add_scaled(s, eRc, Gc, eGc, tm, te);
            add_scaled(st.gm, st.ge, tm, te, st.gm, st.ge);
        } else {
            st.cgv = false;
        }
        // ancestral residue terms (:329-356); the deletion term joins only below an all-gap child
        const bool with_del = cgc && del > 0.0;
        const double sc = with_del ? (eRc < -1000 ? 0.0 : pow2i(eRc)) : 1.0;   // messages are <= 1: eRc <= 0
        // S[p] = sum_x A[x][p] R_c[x]: K independent accumulators, one table row (contiguous in p) per child residue
        double S[K];
#pragma unroll
        for (int p = 0; p < K; p++) S[p] = 0.0;
#pragma unroll
        for (int x = 0; x < K; x++) {
            const double rc = Rc[x];
#pragma unroll
            for (int p = 0; p < K; p++) S[p] = fma(sA[x * K + p], rc, S[p]);
        }
#pragma unroll
        for (int p = 0; p < K; p++) st.Rv[p] *= with_del ? S[p] * sc + del : S[p];
        if (!with_del) st.eRv += eRc;
    }
    // keep the running product's largest entry in [1, 2)
    double mx = st.Rv[0];
#pragma unroll
    for (int p = 1; p < K; p++) mx = fmax(mx, st.Rv[p]);
    if (mx > 0.0) {
        const int e = exponent_of(mx);
        const double f = pow2i(-e);
#pragma unroll
        for (int p = 0; p < K; p++) st.Rv[p] *= f;
        st.eRv += e;
    }
}

// store the node's message; the root also writes the column's log-likelihood
// ent_known / root_known: from the walk descriptor (no dependent lookups on the walker's critical path)
template <int K>
__device__ __forceinline__ void finish_node(const IndelArgs &a, int node, int col, PeelState<K> &st, int ent_known = -1, int root_known = -1)
{
    const size_t nc = (size_t)a.ncols;
    if (st.gm > 0.0) {
        const int e = exponent_of(st.gm);
        st.gm *= pow2i(-e);
        st.ge += e;
    } else {
        st.gm = 0.0;
        st.ge = E_ZERO;
    }
    const int ent = ent_known >= 0 ? ent_known : a.ent_of_node[node];
    double *m = a.msg + ((size_t)ent * (K + 1)) * nc + col;
#pragma unroll
    for (int p = 0; p < K; p++) m[(size_t)p * nc] = st.Rv[p];
    m[(size_t)K * nc] = st.gm;
    a.eR[(size_t)ent * nc + col] = st.eRv;
    a.eG[(size_t)ent * nc + col] = st.ge;
    a.cg[(size_t)ent * nc + col] = st.cgv ? 1 : 0;
    if (root_known >= 0 ? root_known != 0 : a.parent[node] < 0) {   // decorate (:229-252)
        const double LN2 = 0.693147180559945309417232121458;
        double w = 0.0;
#pragma unroll
        for (int r = 0; r < K; r++) w += a.F[r] * st.Rv[r];
        const double resL = w > 0.0 ? (log(w) + (double)st.eRv * LN2) + a.log_geom : -CUDART_INF;
        const double gapL = st.gm > 0.0 ? log(st.gm) + (double)st.ge * LN2 : -CUDART_INF;
        const double mxl = fmax(resL, gapL);
        a.out_logl[col] = mxl == -CUDART_INF ? -CUDART_INF : mxl + log(exp(resL - mxl) + exp(gapL - mxl));
    }
}

template <int K, int THREADS>
__device__ __forceinline__ void stage_table(const IndelArgs &a, int ch, double *sA)
{
    const double *src = a.tb.A + (size_t)ch * (K * K);
    for (int i = threadIdx.x; i < K * K; i += THREADS) sA[i] = src[i];
}

// One height level: grid = (nodes of the level, groups of LEVEL_TPC column tiles).  A CTA stages the tables of its
// node's children ONCE (two at a time: binary nodes never re-stage) and then walks its column tiles.
template <int K>
__global__ void __launch_bounds__(IND_THREADS) indel_level_kernel(IndelArgs a)
{
    __shared__ __align__(16) double sA[2][K * K];
    __shared__ double s_ins[K];
    const int node = a.nodes[blockIdx.x];
    const int c0 = a.first[node], c1 = a.first[node + 1];
    const int tile0 = blockIdx.y * LEVEL_TPC;
    if (threadIdx.x < K) s_ins[threadIdx.x] = a.tb.ins[(size_t)node * K + threadIdx.x];
    if (c1 - c0 <= 2) {
        for (int k = c0; k < c1; k++) stage_table<K, IND_THREADS>(a, a.kids[k], sA[k - c0]);
        __syncthreads();
        for (int tl = 0; tl < LEVEL_TPC; tl++) {
            const int col = (tile0 + tl) * IND_THREADS + threadIdx.x;
            if (col >= a.ncols) break;
            PeelState<K> st;
            st.init();
            for (int k = c0; k < c1; k++) fold_child<K>(a, a.kids[k], col, sA[k - c0], s_ins, st);
            finish_node<K>(a, node, col, st);
        }
        return;
    }
    // any arity: the children's tables pass through shared memory one at a time, per column tile
    for (int tl = 0; tl < LEVEL_TPC; tl++) {
        const int col = (tile0 + tl) * IND_THREADS + threadIdx.x;
        if ((tile0 + tl) * IND_THREADS >= a.ncols) break;
        const bool valid = col < a.ncols;
        PeelState<K> st;
        st.init();
        for (int k = c0; k < c1; k++) {
            __syncthreads();
            stage_table<K, IND_THREADS>(a, a.kids[k], sA[0]);
            __syncthreads();
            if (valid) fold_child<K>(a, a.kids[k], col, sA[0], s_ins, st);
        }
        if (valid) finish_node<K>(a, node, col, st);
    }
}

// The narrow rest of the tree (deep trees: all of it): grid = tiles of WALK_THREADS columns, every CTA walks the node
// list, children before parents.  A 10,000-level caterpillar is one long dependent chain per column, so what a step
// costs is latency: the tables of step i + 1 (its children's A, its own insertion vector) are requested with cp.async
// while step i computes (two shared-memory buffers), and the message a thread wrote in step i stays in its registers
// for step i + 1, whose child it usually is.  Nodes with more than two children stage their tables one by one.
template <int K>
__global__ void __launch_bounds__(WALK_THREADS) indel_walk_kernel(IndelArgs a)
{
    __shared__ __align__(16) double sA[2][2][K * K];
    __shared__ __align__(16) double s_ins[2][K];
    const int col = blockIdx.x * WALK_THREADS + threadIdx.x;
    const bool valid = col < a.ncols;
    const size_t nc = (size_t)a.ncols;
    const int n = a.n_launch_nodes;
    struct Desc { int4 h, c; };   // h = {node, ent, n_children | root << 16, first child index}, c = {child 0, child 1, ent 0, ent 1}
    auto load_desc = [&](int i) -> Desc {
        Desc d;
        if (i < n) { d.h = a.wdesc[2 * i]; d.c = a.wdesc[2 * i + 1]; }
        else { d.h = make_int4(0, -1, 0, 0); d.c = make_int4(-1, -1, -1, -1); }
        return d;
    };
    // tables and insertion vector of the step described by d, into buffer buf (one commit group)
    auto prefetch = [&](const Desc &d, int buf) {
        if (threadIdx.x < K / 2) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_ins[buf][threadIdx.x * 2]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(a.tb.ins + (size_t)d.h.x * K + threadIdx.x * 2) : "memory");
        }
        if ((d.h.z & 0xffff) <= 2)
            for (int k = 0; k < (d.h.z & 0xffff); k++) {
                const double *src = a.tb.A + (size_t)(k == 0 ? d.c.x : d.c.y) * (K * K);
                for (int q = threadIdx.x; q < K * K / 2; q += WALK_THREADS) {
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(&sA[buf][k][q * 2]);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src + q * 2) : "memory");
                }
            }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    // what a thread needs of the step's (at most two) children besides their tables: observed byte of an extant, del
    struct Side { uint8_t ob[2]; double del[2]; };
    auto load_side = [&](const Desc &d) -> Side {
        Side sd;
        sd.ob[0] = sd.ob[1] = BNK_NONE; sd.del[0] = sd.del[1] = 0.0;
        const int nch = d.h.z & 0xffff;
        if (nch <= 2) {
            if (nch > 0) { sd.del[0] = a.tb.del[d.c.x]; if (d.c.z < 0 && valid) sd.ob[0] = a.obs[(size_t)d.c.x * nc + col]; }
            if (nch > 1) { sd.del[1] = a.tb.del[d.c.y]; if (d.c.w < 0 && valid) sd.ob[1] = a.obs[(size_t)d.c.y * nc + col]; }
        }
        return sd;
    };
    if (n <= 0) return;
    Desc d0 = load_desc(0), d1 = load_desc(1);
    prefetch(d0, 0);
    Side s0 = load_side(d0);
    PeelState<K> prev;
    prev.init();
    int prev_ent = -1;
    for (int i = 0; i < n; i++) {
        const int buf = i & 1;
        const Desc d2 = load_desc(i + 2);        // two steps ahead: in registers by the time its tables are requested
        Side s1;
        if (i + 1 < n) {
            prefetch(d1, buf ^ 1);
            s1 = load_side(d1);                  // consumed in the next iteration
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        } else {
            s1 = s0;
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        __syncthreads();
        const int node = d0.h.x, nch = d0.h.z & 0xffff, is_root = d0.h.z >> 16;
        PeelState<K> st;
        st.init();
        if (nch <= 2) {
            if (valid) {
                if (nch > 0) fold_child<K>(a, d0.c.x, col, sA[buf][0], s_ins[buf], st, &prev, prev_ent, d0.c.z, s0.ob[0], s0.del[0]);
                if (nch > 1) fold_child<K>(a, d0.c.y, col, sA[buf][1], s_ins[buf], st, &prev, prev_ent, d0.c.w, s0.ob[1], s0.del[1]);
            }
        } else {
            for (int k = 0; k < nch; k++) {
                if (k > 0) __syncthreads();   // the previous table has been consumed
                stage_table<K, WALK_THREADS>(a, a.kids[d0.h.w + k], sA[buf][0]);
                __syncthreads();
                if (valid) fold_child<K>(a, a.kids[d0.h.w + k], col, sA[buf][0], s_ins[buf], st, &prev, prev_ent);
            }
        }
        if (valid) finish_node<K>(a, node, col, st, d0.h.y, is_root);
        prev = st;
        prev_ent = d0.h.y;
        d0 = d1; d1 = d2; s0 = s1;
        __syncthreads();   // buffer `buf` is free for the prefetch of step i + 2
    }
}

}  // namespace
}  // namespace bnk

// ---- the object behind bnk_indel_* -------------------------------------------------------------------------------
struct bnk_indel {
    BufRegistry reg;
    bnk_model base;
    bnk_tree *tree = nullptr;
    int K = 0, K1 = 0, n = 0, n_int = 0, max_cols = 0, device = 0;
    cudaStream_t stream = nullptr;
    int32_t ncols = -1;          // alignment columns uploaded (the device holds one more: the all-gap column)
    GapModelHost gm;
    bool model_set = false;
    int64_t launches = 0;
    std::vector<std::pair<int32_t, int32_t>> level_ranges;   // into d_nodes, height 1 first
    std::pair<int32_t, int32_t> narrow_range{0, 0};
    DevBuf<int32_t> d_first{&reg}, d_kids{&reg}, d_ent{&reg}, d_parent{&reg}, d_nodes{&reg}, d_eR{&reg}, d_eG{&reg}, d_wdesc{&reg};
    DevBuf<double> d_dist{&reg}, d_gm{&reg}, d_A{&reg}, d_del{&reg}, d_ins{&reg}, d_msg{&reg}, d_logl{&reg}, d_probs{&reg};
    DevBuf<uint8_t> d_obs{&reg}, d_cg{&reg};
    std::vector<double> h_logl;
    // bnk_indel_create_multi: this handle is a front; every part owns one device and a contiguous block of columns
    std::vector<bnk_indel *> parts;
    std::vector<int32_t> part_col0;      // [parts + 1]
    cudaEvent_t ev[3] = {};
    float ms_tables = 0, ms_sweep = 0;
    GapModelDev gmd() const
    {
        const double *b = d_gm.p;
        return GapModelDev{b, b + K1 * K1, b + 2 * K1 * K1, b + 2 * K1 * K1 + K1, gm.mu, gm.lambda, K1};
    }
};

namespace {

template <typename T>
int upload_vec(DevBuf<T> &b, const std::vector<T> &v, cudaStream_t st)
{
    CUDA_TRY(b.ensure(v.size() ? v.size() : 1));
    if (!v.empty()) CUDA_TRY(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return BNK_OK;
}

void indel_free(bnk_indel *h)
{
    if (!h) return;
    for (bnk_indel *p : h->parts) indel_free(p);
    h->parts.clear();
    if (!h->stream && h->reg.all.empty() && !h->tree) { delete h; return; }
    DeviceGuard guard(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    for (auto &e : h->ev) if (e) cudaEventDestroy(e);
    for (DevBufBase *b : h->reg.all) b->release();
    if (h->tree) tree_release(h->tree);
    delete h;
}

// ksiT on the host (GapSubstModel.ksiT :184-192), for probExtraCol
double ksi_host(const GapModelHost &g, double t)
{
    if (g.mu == 0 && g.lambda == 0) return 0.0;
    const double insertionRate = g.lambda / (g.mu + g.lambda);
    const double indelProp = 1 - std::exp(-((g.mu + g.lambda) * t));
    return insertionRate * indelProp;
}

// MathEx.logm1exp (src/smile/math/MathEx.java:4573-4584)
int logm1exp(double x, double *out)
{
    const double LOG_HALF = std::log(0.5);
    if (x > 0) return fail(BNK_ERR_ARG, "indel: log P(all-gap column) = %g > 0", x);
    *out = x <= LOG_HALF ? std::log1p(-std::exp(x)) : std::log(-std::expm1(x));
    return BNK_OK;
}

// tables for `rate`, then the sweep over the uploaded columns + the all-gap column; leaves the column
// log-likelihoods in h->h_logl (ncols + 1 values)
int run_sweep(bnk_indel *h, double rate, double geom);

// the front of a multi-device handle: every part sweeps its column block on its own device at the same time (one
// host thread each; the parts share nothing), then the column values are laid end to end -- so every sum over
// columns downstream runs in column order whatever the number of devices (the reduction of
// IndelPeeler.calcProbAlnGivenTree, src/asr/IndelPeeler.java:131-136; 8 bytes per column come back anyway)
int run_sweep_multi(bnk_indel *h, double rate, double geom)
{
    if (h->ncols < 0) return fail(BNK_ERR_STATE, "indel: no alignment uploaded (bnk_indel_set_alignment)");
    const size_t np = h->parts.size();
    std::vector<int> rcs(np, BNK_OK);
    std::vector<std::string> msgs(np);
    std::vector<std::thread> th;
    for (size_t i = 0; i < np; i++)
        th.emplace_back([&, i] {
            bnk_indel *p = h->parts[i];
            if (p->ncols < 0) return;   // more devices than columns
            DeviceGuard guard(p->device);
            rcs[i] = guard.err == cudaSuccess ? run_sweep(p, rate, geom) : BNK_ERR_CUDA;
            if (rcs[i] != BNK_OK) msgs[i] = bnk_last_error();
        });
    for (auto &t : th) t.join();
    for (size_t i = 0; i < np; i++)
        if (rcs[i] != BNK_OK) return fail(rcs[i], "indel (device %d): %s", h->parts[i]->device, msgs[i].c_str());
    h->h_logl.assign((size_t)h->ncols + 1, 0.0);
    h->ms_tables = 0; h->ms_sweep = 0; h->launches = 0;
    for (size_t i = 0; i < np; i++) {
        bnk_indel *p = h->parts[i];
        if (p->ncols < 0) continue;
        std::memcpy(h->h_logl.data() + h->part_col0[i], p->h_logl.data(), sizeof(double) * p->ncols);
        if (i == 0) h->h_logl[h->ncols] = p->h_logl[p->ncols];   // the all-gap column: the same on every device
        h->ms_tables = std::max(h->ms_tables, p->ms_tables);
        h->ms_sweep = std::max(h->ms_sweep, p->ms_sweep);
        h->launches += p->launches;
    }
    return BNK_OK;
}

int run_sweep(bnk_indel *h, double rate, double geom)
{
    if (!h->parts.empty()) return run_sweep_multi(h, rate, geom);
    if (h->ncols < 0) return fail(BNK_ERR_STATE, "indel: no alignment uploaded (bnk_indel_set_alignment)");
    if (!h->model_set) return fail(BNK_ERR_STATE, "indel: no indel rates set (bnk_indel_set_rates)");
    if (!(geom > 0.0)) return fail(BNK_ERR_ARG, "indel: the geometric sequence-length parameter must be > 0");
    const int nc = h->ncols + 1, K = h->K;
    cudaStream_t st = h->stream;
    IndelTables tb{h->d_A.p, h->d_del.p, h->d_ins.p};
    CUDA_TRY(cudaEventRecord(h->ev[0], st));
    launch_indel_tables(h->K1, h->gmd(), h->d_dist.p, h->d_parent.p, rate, h->n, tb, st);
    h->launches++;
    CUDA_TRY(cudaEventRecord(h->ev[1], st));
    IndelArgs a{};
    a.first = h->d_first.p; a.kids = h->d_kids.p; a.ent_of_node = h->d_ent.p; a.parent = h->d_parent.p;
    a.obs = h->d_obs.p; a.ncols = nc; a.tb = tb; a.F = h->d_gm.p + 2 * h->K1 * h->K1 + h->K1;
    a.msg = h->d_msg.p; a.eR = h->d_eR.p; a.eG = h->d_eG.p; a.cg = h->d_cg.p;
    a.log_geom = std::log(geom); a.out_logl = h->d_logl.p;
    const int tiles = (nc + IND_THREADS - 1) / IND_THREADS;
    const int tile_groups = (tiles + LEVEL_TPC - 1) / LEVEL_TPC;
    for (const auto &lr : h->level_ranges) {
        a.nodes = h->d_nodes.p + lr.first; a.n_launch_nodes = lr.second - lr.first;
        dim3 grid(a.n_launch_nodes, tile_groups);
        if (K == 20) indel_level_kernel<20><<<grid, IND_THREADS, 0, st>>>(a);
        else indel_level_kernel<4><<<grid, IND_THREADS, 0, st>>>(a);
        h->launches++;
    }
    if (h->narrow_range.second > h->narrow_range.first) {
        a.nodes = h->d_nodes.p + h->narrow_range.first; a.n_launch_nodes = h->narrow_range.second - h->narrow_range.first;
        a.wdesc = reinterpret_cast<const int4 *>(h->d_wdesc.p);
        const int wtiles = (nc + WALK_THREADS - 1) / WALK_THREADS;
        if (K == 20) indel_walk_kernel<20><<<wtiles, WALK_THREADS, 0, st>>>(a);
        else indel_walk_kernel<4><<<wtiles, WALK_THREADS, 0, st>>>(a);
        h->launches++;
    }
    CUDA_TRY(cudaEventRecord(h->ev[2], st));
    CUDA_TRY(cudaGetLastError());
    h->h_logl.resize(nc);
    CUDA_TRY(cudaMemcpyAsync(h->h_logl.data(), h->d_logl.p, sizeof(double) * nc, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&h->ms_tables, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&h->ms_sweep, h->ev[1], h->ev[2]);
    return BNK_OK;
}

// IndelPeeler.calcProbAlnGivenTree (:125-161) at the current (mu, lambda)
int aln_logl(bnk_indel *h, double geom, double *out)
{
    TRY(run_sweep(h, 1.0, geom));
    double logLikelihood = 0.0;
    for (int c = 0; c < h->ncols; c++) logLikelihood += h->h_logl[c];
    // probExtraCol (:163-193): tree distances without a rate
    const bnk_tree *t = h->tree;
    std::vector<double> pStar(h->n, 0.0);
    for (int v = h->n - 1; v >= 0; v--) {
        if (t->first[v + 1] == t->first[v]) { pStar[v] = 0.0; continue; }
        double childrenLL = 0.0;
        for (int k = t->first[v]; k < t->first[v + 1]; k++) {
            const int u = t->kids[k];
            childrenLL += pStar[u];
            childrenLL += std::log(1 - ksi_host(h->gm, t->dist[u]));
        }
        pStar[v] = childrenLL;
    }
    double extra = 0.0;
    extra += std::log(1 - geom);
    extra += pStar[0];
    double unobserved = 0.0;
    TRY(logm1exp(h->h_logl[h->ncols], &unobserved));
    *out = (extra - unobserved) + logLikelihood;
    return BNK_OK;
}

int set_rates(bnk_indel *h, double mu, double lambda)
{
    if (!h->parts.empty()) {
        TRY(gap_model_from_base(&h->base, mu, lambda, &h->gm));   // the front keeps the host model (probExtraCol)
        for (bnk_indel *p : h->parts) {
            DeviceGuard guard(p->device);
            if (guard.err != cudaSuccess) return fail(BNK_ERR_CUDA, "indel: cudaSetDevice(%d)", p->device);
            TRY(set_rates(p, mu, lambda));
        }
        h->model_set = true;
        return BNK_OK;
    }
    TRY(gap_model_from_base(&h->base, mu, lambda, &h->gm));
    const int K1 = h->K1;
    std::vector<double> pack((size_t)2 * K1 * K1 + 2 * K1);
    std::memcpy(pack.data(), h->gm.evec, sizeof(double) * K1 * K1);
    std::memcpy(pack.data() + K1 * K1, h->gm.ievec, sizeof(double) * K1 * K1);
    std::memcpy(pack.data() + 2 * K1 * K1, h->gm.eval, sizeof(double) * K1);
    std::memcpy(pack.data() + 2 * K1 * K1 + K1, h->gm.F, sizeof(double) * K1);
    TRY(upload_vec(h->d_gm, pack, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));   // `pack` goes out of scope
    h->model_set = true;
    return BNK_OK;
}

}  // namespace

extern "C" int bnk_indel_create(const bnk_model *base, const bnk_tree *tree, int32_t max_cols, int device, bnk_indel **out)
{
    if (!base || !tree || !out) return fail(BNK_ERR_ARG, "indel: null argument");
    if (base->K != 20 && base->K != 4) return fail(BNK_ERR_ARG, "indel: 20- or 4-state base models only");
    if (!base->has_R) return fail(BNK_ERR_MODEL, "indel: the base model has no rate matrix (adopted eigensystem)");
    if (max_cols < 1) return fail(BNK_ERR_ARG, "indel: max_cols must be >= 1");
    if (tree->root_count != 1 || tree->parent[0] >= 0) return fail(BNK_ERR_TREE, "indel: one tree rooted at node 0 expected");
    if (tree->n_int < 1) return fail(BNK_ERR_TREE, "indel: the tree has no ancestor");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return fail(BNK_ERR_CUDA, "indel: no CUDA device"); }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= ndev) return fail(BNK_ERR_ARG, "indel: device %d of %d", device, ndev);
    bnk_indel *h = new (std::nothrow) bnk_indel();
    if (!h) return fail(BNK_ERR_NOMEM, "indel: out of host memory");
    h->base = *base;
    h->tree = const_cast<bnk_tree *>(tree);
    tree_retain(h->tree);
    h->K = base->K; h->K1 = base->K + 1; h->n = tree->n; h->n_int = tree->n_int; h->max_cols = max_cols; h->device = device;
    auto bail = [&](int rc) { indel_free(h); return rc; };
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return bail(fail(BNK_ERR_CUDA, "indel: cudaSetDevice: %s", cudaGetErrorString(guard.err)));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10)
        return bail(fail(BNK_ERR_CUDA, "indel: device %d is not an sm_100-class GPU", device));
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { h->stream = nullptr; return bail(fail(BNK_ERR_CUDA, "indel: cudaStreamCreate")); }
    for (auto &e : h->ev) if (cudaEventCreate(&e) != cudaSuccess) return bail(fail(BNK_ERR_CUDA, "indel: cudaEventCreate"));
    // schedule: height levels that are wide enough, then the rest in decreasing index order
    int hmax = 0;
    for (int v = 0; v < h->n; v++) hmax = std::max(hmax, tree->height[v]);
    std::vector<int32_t> width(hmax + 2, 0);
    for (int v = 0; v < h->n; v++) width[tree->height[v]]++;
    int wide_h = 0;
    while (wide_h + 1 <= hmax && width[wide_h + 1] >= LEVEL_MIN_NODES) wide_h++;
    std::vector<int32_t> nodes;
    for (int lv = 1; lv <= wide_h; lv++) {
        const int32_t lo = (int32_t)nodes.size();
        for (int v = 0; v < h->n; v++) if (tree->height[v] == lv) nodes.push_back(v);
        h->level_ranges.push_back({lo, (int32_t)nodes.size()});
    }
    {
        const int32_t lo = (int32_t)nodes.size();
        for (int v = h->n - 1; v >= 0; v--) if (tree->height[v] > wide_h) nodes.push_back(v);
        h->narrow_range = {lo, (int32_t)nodes.size()};
    }
    auto step = [&](int rc) { return rc; };
    int rc = BNK_OK;
    if ((rc = step(upload_vec(h->d_first, tree->first, h->stream))) != BNK_OK) return bail(rc);
    if ((rc = step(upload_vec(h->d_kids, tree->kids, h->stream))) != BNK_OK) return bail(rc);
    if ((rc = step(upload_vec(h->d_ent, tree->ent_of_node, h->stream))) != BNK_OK) return bail(rc);
    if ((rc = step(upload_vec(h->d_parent, tree->parent, h->stream))) != BNK_OK) return bail(rc);
    if ((rc = step(upload_vec(h->d_dist, tree->dist, h->stream))) != BNK_OK) return bail(rc);
    if ((rc = step(upload_vec(h->d_nodes, nodes, h->stream))) != BNK_OK) return bail(rc);
    {   // walk descriptors of the narrow nodes: one 32-byte record per step instead of four dependent lookups
        std::vector<int32_t> wd;
        for (int32_t i = h->narrow_range.first; i < h->narrow_range.second; i++) {
            const int v = nodes[i], c0 = tree->first[v], nch = tree->first[v + 1] - c0;
            const int ch0 = nch > 0 ? tree->kids[c0] : -1, ch1 = nch > 1 ? tree->kids[c0 + 1] : -1;
            if (nch > 0xffff) return bail(fail(BNK_ERR_TREE, "indel: a narrow node with %d children", nch));
            const int32_t rec[8] = {v, tree->ent_of_node[v], nch | (tree->parent[v] < 0 ? 1 << 16 : 0), c0, ch0, ch1,
                                    ch0 >= 0 ? tree->ent_of_node[ch0] : -1, ch1 >= 0 ? tree->ent_of_node[ch1] : -1};
            wd.insert(wd.end(), rec, rec + 8);
        }
        if (wd.empty()) wd.assign(8, 0);
        if ((rc = step(upload_vec(h->d_wdesc, wd, h->stream))) != BNK_OK) return bail(rc);
    }
    const size_t nc = (size_t)max_cols + 1, K = h->K;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = h->d_A.ensure((size_t)h->n * K * K);
    if (e == cudaSuccess) e = h->d_del.ensure(h->n);
    if (e == cudaSuccess) e = h->d_ins.ensure((size_t)h->n * K);
    if (e == cudaSuccess) e = h->d_obs.ensure((size_t)h->n * nc);
    if (e == cudaSuccess) e = h->d_msg.ensure((size_t)h->n_int * (K + 1) * nc);
    if (e == cudaSuccess) e = h->d_eR.ensure((size_t)h->n_int * nc);
    if (e == cudaSuccess) e = h->d_eG.ensure((size_t)h->n_int * nc);
    if (e == cudaSuccess) e = h->d_cg.ensure((size_t)h->n_int * nc);
    if (e == cudaSuccess) e = h->d_logl.ensure(nc);
    if (e == cudaSuccess) e = h->d_probs.ensure((size_t)h->K1 * h->K1);
    if (e != cudaSuccess) return bail(fail(e == cudaErrorMemoryAllocation ? BNK_ERR_NOMEM : BNK_ERR_CUDA, "indel: device allocation: %s", cudaGetErrorString(e)));
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return bail(fail(BNK_ERR_CUDA, "indel: upload failed"));
    *out = h;
    return BNK_OK;
}

// All devices behind one handle (the GPU counterpart of GRASP.NTHREADS in runPeelingJobs, IndelPeeler.java:208-224):
// contiguous column blocks, one per device; n_devices <= 0 = every visible device
extern "C" int bnk_indel_create_multi(const bnk_model *base, const bnk_tree *tree, int32_t max_cols, int n_devices,
                                      const int *devices_or_null, bnk_indel **out)
{
    if (!base || !tree || !out) return fail(BNK_ERR_ARG, "indel: null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return fail(BNK_ERR_CUDA, "indel: no CUDA device"); }
    if (n_devices <= 0) n_devices = ndev;
    if (!devices_or_null && n_devices > ndev) return fail(BNK_ERR_ARG, "indel: %d devices asked for, %d visible", n_devices, ndev);
    bnk_indel *h = new (std::nothrow) bnk_indel();
    if (!h) return fail(BNK_ERR_NOMEM, "indel: out of host memory");
    h->base = *base;
    h->K = base->K; h->K1 = base->K + 1; h->n = tree->n; h->n_int = tree->n_int; h->max_cols = max_cols;
    const int32_t per = (max_cols + n_devices - 1) / n_devices;
    for (int i = 0; i < n_devices; i++) {
        bnk_indel *p = nullptr;
        const int rc = bnk_indel_create(base, tree, std::max(per, 1), devices_or_null ? devices_or_null[i] : i, &p);
        if (rc != BNK_OK) { indel_free(h); return rc; }
        h->parts.push_back(p);
    }
    h->device = h->parts[0]->device;
    h->tree = h->parts[0]->tree;   // probExtraCol reads the tree; the parts hold the references
    *out = h;
    return BNK_OK;
}

extern "C" void bnk_indel_destroy(bnk_indel *h)
{
    if (h && !h->parts.empty()) h->tree = nullptr;   // a front borrows its tree pointer from part 0
    indel_free(h);
}

extern "C" int bnk_indel_set_alignment(bnk_indel *h, int32_t n_cols, const uint8_t *obs)
{
    if (!h || !obs) return fail(BNK_ERR_ARG, "indel: null argument");
    if (n_cols < 1 || n_cols > h->max_cols) return fail(BNK_ERR_ARG, "indel: n_cols %d outside 1..%d", n_cols, h->max_cols);
    if (!h->parts.empty()) {
        // contiguous blocks whose sizes differ by at most one (the rule of bnk_mgpu_shard / dist.shard_columns)
        const int np = (int)h->parts.size();
        h->part_col0.assign(np + 1, 0);
        std::vector<uint8_t> block;
        for (int i = 0; i < np; i++) {
            const int32_t base_w = n_cols / np, rem = n_cols % np;
            const int32_t lo = i * base_w + std::min(i, rem), w = base_w + (i < rem ? 1 : 0);
            h->part_col0[i] = lo; h->part_col0[i + 1] = lo + w;
            bnk_indel *p = h->parts[i];
            p->ncols = -1;
            if (w == 0) continue;
            block.resize((size_t)h->n * w);
            for (int v = 0; v < h->n; v++) std::memcpy(block.data() + (size_t)v * w, obs + (size_t)v * n_cols + lo, w);
            TRY(bnk_indel_set_alignment(p, w, block.data()));
        }
        h->ncols = n_cols;
        return BNK_OK;
    }
    ENTER_DEVICE(h->device);
    const bnk_tree *t = h->tree;
    for (int v = 0; v < h->n; v++) {   // validate the extants' rows (a byte >= K would index past a table)
        if (t->first[v + 1] != t->first[v]) continue;
        const uint8_t *row = obs + (size_t)v * n_cols;
        for (int c = 0; c < n_cols; c++)
            if (row[c] != BNK_NONE && row[c] >= h->K) return fail(BNK_ERR_ARG, "indel: state %d at node %d, column %d (alphabet of %d)", row[c], v, c, h->K);
    }
    const size_t nc = (size_t)n_cols + 1;
    // device layout [node][n_cols + 1]: the extra column is the all-gap column of calcProbAlnGivenTree (:139-158)
    CUDA_TRY(cudaMemsetAsync(h->d_obs.p, BNK_NONE, (size_t)h->n * nc, h->stream));
    CUDA_TRY(cudaMemcpy2DAsync(h->d_obs.p, nc, obs, n_cols, n_cols, h->n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->ncols = n_cols;
    return BNK_OK;
}

extern "C" int bnk_indel_set_rates(bnk_indel *h, double mu, double lambda)
{
    if (!h) return fail(BNK_ERR_ARG, "indel: null argument");
    ENTER_DEVICE(h->device);
    return set_rates(h, mu, lambda);
}

extern "C" int bnk_indel_column_logl(bnk_indel *h, double rate, double geom, double *out_logl, double *out_sum_or_null)
{
    if (!h || !out_logl) return fail(BNK_ERR_ARG, "indel: null argument");
    ENTER_DEVICE(h->device);
    TRY(run_sweep(h, rate, geom));
    std::memcpy(out_logl, h->h_logl.data(), sizeof(double) * h->ncols);
    if (out_sum_or_null) {
        double s = 0.0;
        for (int c = 0; c < h->ncols; c++) s += h->h_logl[c];
        *out_sum_or_null = s;
    }
    return BNK_OK;
}

extern "C" int bnk_indel_column_priors(bnk_indel *h, double geom, int32_t n_rates, const double *rates, double *out)
{
    if (!h || !rates || !out || n_rates < 1) return fail(BNK_ERR_ARG, "indel: invalid argument");
    ENTER_DEVICE(h->device);
    for (int r = 0; r < n_rates; r++) {
        TRY(run_sweep(h, rates[r], geom));
        for (int c = 0; c < h->ncols; c++) out[(size_t)c * n_rates + r] = h->h_logl[c];   // columnPriors[colIdx][rateIdx]
    }
    return BNK_OK;
}

extern "C" int bnk_indel_aln_logl(bnk_indel *h, double geom, double *out)
{
    if (!h || !out) return fail(BNK_ERR_ARG, "indel: null argument");
    ENTER_DEVICE(h->device);
    return aln_logl(h, geom, out);
}

// Minimise.brent (src/smile/math/special/Minimise.java:25-107) over mu = lambda, f = -calcProbAlnGivenTree; the
// reference evaluates f(v), f(w), f(x) anew in every recursion -- f is deterministic, the values are kept here
extern "C" int bnk_indel_optimise_mulambda(bnk_indel *h, double min_val, double max_val, double geom, double *out_mulambda,
                                           int32_t *n_evals_or_null)
{
    if (!h || !out_mulambda) return fail(BNK_ERR_ARG, "indel: null argument");
    if (!(min_val >= 0.0) || !(max_val > min_val)) return fail(BNK_ERR_ARG, "indel: invalid search interval");
    ENTER_DEVICE(h->device);
    const int MAX = 500;
    const double EPS = 10e-6;
    const double INVPHI = (1.0 + std::sqrt(5.0)) / 2.0 - 1.0;
    const double RATIO = (3.0 - std::sqrt(5.0)) / 2;
    std::vector<std::pair<double, double>> seen;
    int status = BNK_OK;
    auto f = [&](double x) -> double {
        for (const auto &s : seen) if (s.first == x) return s.second;
        double v = 0.0;
        if (status == BNK_OK) status = set_rates(h, x, x);
        if (status == BNK_OK) status = aln_logl(h, geom, &v);
        seen.push_back({x, -v});
        return -v;
    };
    double a = min_val, b = max_val, x = b + INVPHI * (a - b), v = x, w = x, dold = 0, eold = 0, m = 0;
    for (int i = 0;; i++) {
        const double fv = f(v), fw = f(w), fx = f(x);
        if (status != BNK_OK) return status;
        m = 0.5 * (a + b);
        if (b - a <= EPS || i > MAX) break;
        const double r = (x - w) * (fx - fv), tq = (x - v) * (fx - fw), tp = (x - v) * tq - (x - w) * r, tq2 = 2.0 * (tq - r);
        const double p = tq2 > 0.0 ? -tp : tp, q = tq2 > 0.0 ? tq2 : -tq2;
        const bool safe = q != 0.0;
        const double deltax = safe ? p / q : 0.0;
        const bool parabolic = safe && a < x + deltax && x + deltax < b && std::fabs(deltax) < 0.5 * std::fabs(eold);
        const double e = parabolic ? dold : (x < m ? b - x : a - x);
        const double d = parabolic ? deltax : RATIO * e;
        const double u = x + d;
        const double fu = f(u);
        if (status != BNK_OK) return status;
        if (fu <= fx) {
            if (u < x) b = x; else a = x;
            v = w; w = x; x = u;
        } else {
            if (u < x) a = u; else b = u;
            if (fu <= fw || w == x) { v = w; w = u; }
            else if (fu <= fv || v == x || v == w) { v = u; }
        }
        dold = d; eold = e;
    }
    *out_mulambda = m;
    if (n_evals_or_null) *n_evals_or_null = (int32_t)seen.size();
    return BNK_OK;
}

extern "C" int bnk_indel_probs(bnk_indel *h, double t, double *out_P, double *out_ksi_gamma_or_null, double *out_F_or_null)
{
    if (!h || !out_P) return fail(BNK_ERR_ARG, "indel: null argument");
    if (!h->model_set) return fail(BNK_ERR_STATE, "indel: no indel rates set (bnk_indel_set_rates)");
    ENTER_DEVICE(h->device);
    launch_indel_probs(h->K1, h->gmd(), t, h->d_probs.p, h->stream);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_P, h->d_probs.p, sizeof(double) * h->K1 * h->K1, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (out_ksi_gamma_or_null) {
        const double ksi = ksi_host(h->gm, t);
        double gamma = 0.0;
        if (!(h->gm.mu == 0 && h->gm.lambda == 0)) {
            const double total = h->gm.mu + h->gm.lambda;
            gamma = (h->gm.mu / total) * (1 - std::exp(-(total * t)));
        }
        out_ksi_gamma_or_null[0] = ksi;
        out_ksi_gamma_or_null[1] = gamma;
    }
    if (out_F_or_null) std::memcpy(out_F_or_null, h->gm.F, sizeof(double) * h->K1);
    return BNK_OK;
}

extern "C" double bnk_indel_phase_ms(const bnk_indel *h, int phase)
{
    if (!h) return -1.0;
    return phase == 0 ? h->ms_tables : h->ms_sweep;
}

extern "C" int64_t bnk_indel_launch_count(const bnk_indel *h) { return h ? h->launches : 0; }
extern "C" int64_t bnk_indel_device_bytes(const bnk_indel *h) { return h ? h->reg.bytes : 0; }
