// multi.cu -- (1) the pipelined host-buffer path: a call over host arrays is cut into column chunks that run
// through two sub-workspaces on their own streams, so the result copy of one chunk (8 GB of posteriors at
// BASELINE's C3) overlaps the kernels and the input copy of the next; (2) bnk_mgpu_*: the same column ranges
// spread over several devices of one process -- the device-side counterpart of the reference's thread pool
// over columns (ThreadedDecorators, src/asr/ThreadedDecorators.java:28-72; GRASP.NTHREADS, src/asr/GRASP.java:35).
//
// Columns are independent (one TreeDecor job per column in the reference), so neither level needs a data-path
// collective.  The one exchange is the sum of the per-device column log-likelihoods of an optimisation loop
// (the reduction of IndelPeeler.calcProbAlnGivenTree, src/asr/IndelPeeler.java:125-150): ncclAllReduce over the
// devices' streams when libnccl can be loaded, else the per-device sums are added on the host.
#include "workspace.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <thread>

using namespace bnk;

// --------------------------------------------------------------------------------------
// column-chunk pipeline on one device
// --------------------------------------------------------------------------------------
namespace bnk {

bool ws_wants_pipeline(const bnk_ws *ws, int32_t n_cols)
{
    if (ws->is_sub || ws->pipeline_cols < 0) return false;
    if (ws->pipeline_cols > 0) return n_cols > ws->pipeline_cols;
    // automatic: only where the copies are worth hiding (the chunks pay a table build each)
    return n_cols >= 2048 && (size_t)ws->n * n_cols >= ((size_t)32 << 20);
}

static int32_t chunk_cols_for(const bnk_ws *ws, int32_t ncols)
{
    if (ws->pipeline_cols > 0) return std::min(ncols, ws->pipeline_cols);
    if (ncols < 2048) return ncols;
    const int32_t quarter = (ncols + 3) / 4;
    return std::min(ncols, std::max(512, (quarter + 127) / 128 * 128));
}

static int ensure_subs(bnk_ws *ws, int32_t chunk)
{
    if (ws->subs.size() == 2 && ws->sub_cols >= chunk) return BNK_OK;
    for (bnk_ws *s : ws->subs) ws_free(s);
    ws->subs.clear();
    for (int k = 0; k < 2; k++) {
        bnk_ws *sub = nullptr;
        TRY(bnk_ws_create(&ws->model, ws->tree, chunk, ws->device, nullptr, &sub));
        sub->is_sub = true;
        ws->subs.push_back(sub);
    }
    ws->sub_cols = chunk;
    ws->sub_dist_version = 0;
    if (!ws->copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ws->copy_stream, cudaStreamNonBlocking));
        for (auto &e : ws->chunk_done) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto &e : ws->copy_done) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    return BNK_OK;
}

int ws_run_host_range(bnk_ws *ws, const HostCall &hc)
{
    ENTER_DEVICE(ws->device);
    if (hc.ncols < 1) return BNK_OK;
    const int32_t chunk = chunk_cols_for(ws, hc.ncols);
    TRY(ensure_subs(ws, chunk));
    for (bnk_ws *sub : ws->subs) {  // the sub-workspaces follow the parent's settings
        sub->cfg_tile_cols = ws->cfg_tile_cols; sub->cfg_cpt = ws->cfg_cpt; sub->force_generic = ws->force_generic;
        sub->no_wide = ws->no_wide; sub->evidence_mode = ws->evidence_mode;
        if (ws->sub_dist_version != ws->dist_version + 1) TRY(bnk_set_distances(sub, ws->dist.data()));
    }
    ws->sub_dist_version = ws->dist_version + 1;
    const int K = ws->K, n = ws->n;
    const size_t stride = (size_t)hc.n_cols_total;
    const int n_chunks = (hc.ncols + chunk - 1) / chunk;
    std::vector<double> partial((size_t)n_chunks, 0.0);
    const int32_t nq = hc.n_query < 0 ? ws->n_int : hc.n_query;
    int rc = BNK_OK;
    for (int j = 0; j < n_chunks && rc == BNK_OK; j++) {
        bnk_ws *w = ws->subs[j & 1];
        // even split: no ragged last chunk
        const int32_t c0 = hc.col0 + (int32_t)((int64_t)hc.ncols * j / n_chunks);
        const int32_t nc = hc.col0 + (int32_t)((int64_t)hc.ncols * (j + 1) / n_chunks) - c0;
        auto body = [&]() -> int {
            // Results leave on ONE copy stream: chunk j's kernels (stream j & 1) -> event -> its D2H queued behind chunk
            // j - 1's.  The buffers of this sub-workspace are free again when the copy of chunk j - 2 is done; its
            // stream waits for that event, the host does not (r2 first cut: both D2H copies ran concurrently on the
            // two streams, shared PCIe, finished together and left the link idle while the next chunk computed).
            TRY(bnk_set_rates(w, nc, hc.rates ? hc.rates + c0 : nullptr));
            if (j >= 2) CUDA_TRY(cudaStreamWaitEvent(w->stream, ws->copy_done[j & 1], 0));
            const size_t bytes = (size_t)n * nc;
            CUDA_TRY(w->d_obs_in.ensure(bytes));
            CUDA_TRY(cudaMemcpy2DAsync(w->d_obs_in.p, nc, hc.obs + c0, stride, nc, n, cudaMemcpyHostToDevice, w->stream));
            const uint8_t *d_present = nullptr;
            if (hc.present) {
                CUDA_TRY(w->d_present_in.ensure(bytes));
                CUDA_TRY(cudaMemcpy2DAsync(w->d_present_in.p, nc, hc.present + c0, stride, nc, n, cudaMemcpyHostToDevice, w->stream));
                d_present = w->d_present_in.p;
            }
            if (hc.op == 0) {
                CUDA_TRY(w->d_state.ensure(bytes));
                TRY(bnk_joint_dev(w, nc, w->d_obs_in.p, d_present, w->d_state.p));
                CUDA_TRY(cudaEventRecord(ws->chunk_done[j & 1], w->stream));
                CUDA_TRY(cudaStreamWaitEvent(ws->copy_stream, ws->chunk_done[j & 1], 0));
                CUDA_TRY(cudaMemcpy2DAsync(hc.out_state + c0, stride, w->d_state.p, nc, nc, n, cudaMemcpyDeviceToHost, ws->copy_stream));
                CUDA_TRY(cudaEventRecord(ws->copy_done[j & 1], ws->copy_stream));
            } else if (hc.op == 1 || hc.op == 3) {
                const bool want_post = hc.out_post && nq > 0;
                if (want_post) CUDA_TRY(w->d_post.ensure((size_t)nq * nc * K));
                if (hc.out_logl) CUDA_TRY(w->d_logl.ensure(nc));
                if (hc.op == 3) {  // joint + marginal of the chunk: the joint pass hides under the previous chunk's result copy
                    CUDA_TRY(w->d_state.ensure(bytes));
                    TRY(bnk_joint_dev(w, nc, w->d_obs_in.p, d_present, w->d_state.p));
                }
                TRY(ws_run_sum(w, nc, w->d_obs_in.p, d_present, hc.query_nodes, hc.n_query, want_post ? w->d_post.p : nullptr,
                               hc.out_logl ? w->d_logl.p : nullptr, nullptr));
                CUDA_TRY(cudaEventRecord(ws->chunk_done[j & 1], w->stream));
                CUDA_TRY(cudaStreamWaitEvent(ws->copy_stream, ws->chunk_done[j & 1], 0));
                if (hc.op == 3)
                    CUDA_TRY(cudaMemcpy2DAsync(hc.out_state + c0, stride, w->d_state.p, nc, nc, n, cudaMemcpyDeviceToHost, ws->copy_stream));
                if (want_post)
                    CUDA_TRY(cudaMemcpy2DAsync(hc.out_post + (size_t)c0 * K, stride * K * sizeof(double), w->d_post.p,
                                               (size_t)nc * K * sizeof(double), (size_t)nc * K * sizeof(double), nq,
                                               cudaMemcpyDeviceToHost, ws->copy_stream));
                if (hc.out_logl)
                    CUDA_TRY(cudaMemcpyAsync(hc.out_logl + c0, w->d_logl.p, sizeof(double) * nc, cudaMemcpyDeviceToHost, ws->copy_stream));
                CUDA_TRY(cudaEventRecord(ws->copy_done[j & 1], ws->copy_stream));
            } else {
                CUDA_TRY(w->d_logl.ensure(nc));
                TRY(ws_run_sum(w, nc, w->d_obs_in.p, d_present, nullptr, 0, nullptr, w->d_logl.p, w->d_sum.p));
                if (hc.out_logl)
                    CUDA_TRY(cudaMemcpyAsync(hc.out_logl + c0, w->d_logl.p, sizeof(double) * nc, cudaMemcpyDeviceToHost, w->stream));
                CUDA_TRY(cudaMemcpyAsync(&partial[j], w->d_sum.p, sizeof(double), cudaMemcpyDeviceToHost, w->stream));
            }
            return BNK_OK;
        };
        rc = body();
    }
    for (bnk_ws *sub : ws->subs) {
        const cudaError_t e = cudaStreamSynchronize(sub->stream);
        if (e != cudaSuccess && rc == BNK_OK) rc = fail(BNK_ERR_CUDA, "pipeline: %s", cudaGetErrorString(e));
    }
    {
        const cudaError_t e = cudaStreamSynchronize(ws->copy_stream);
        if (e != cudaSuccess && rc == BNK_OK) rc = fail(BNK_ERR_CUDA, "pipeline copies: %s", cudaGetErrorString(e));
    }
    if (rc == BNK_OK && hc.out_sum)
        for (double p : partial) *hc.out_sum += p;
    return rc;
}

}  // namespace bnk

// --------------------------------------------------------------------------------------
// several devices, one process
// --------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    bool load()
    {
        if (lib) return true;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        if (!CommInitAll || !CommDestroy || !AllReduce || !GroupStart || !GroupEnd) { dlclose(lib); lib = nullptr; }
        return lib != nullptr;
    }
};
}  // namespace

struct bnk_mgpu {
    std::vector<bnk_ws *> ws;
    std::vector<int> dev;
    int32_t max_cols = 0;
    std::vector<double> rates;  // per column of the whole alignment (empty: 1.0)
    // resident inputs of an optimisation loop
    int32_t res_cols = -1;
    bool res_present = false;
    // the reduction
    NcclApi nccl;
    std::vector<ncclComm_t> comms;
    bool nccl_tried = false;
    std::string reduce_how = "none yet";
};

static void shard_of(int n_dev, int32_t n_cols, int idx, int32_t &c0, int32_t &nc)
{
    // contiguous blocks whose sizes differ by at most one column (bnkit_b200/dist.py: shard_columns)
    const int32_t base = n_cols / n_dev, rem = n_cols % n_dev;
    c0 = idx * base + std::min(idx, rem);
    nc = base + (idx < rem ? 1 : 0);
}

extern "C" int bnk_mgpu_create(const bnk_model *m, const bnk_tree *t, int32_t max_cols, int n_devices, const int *devices,
                               bnk_mgpu **out)
{
    if (!m || !t || !out || max_cols < 1) return fail(BNK_ERR_ARG, "mgpu: bad argument");
    const int visible = bnk_device_count();
    if (visible < 1) return fail(BNK_ERR_CUDA, "mgpu: no sm_100 device visible (libbnkgpu has no CPU fallback)");
    if (n_devices <= 0) n_devices = visible;
    if (!devices && n_devices > visible) return fail(BNK_ERR_ARG, "mgpu: %d devices asked for, %d visible", n_devices, visible);
    bnk_mgpu *g = new (std::nothrow) bnk_mgpu();
    if (!g) return fail(BNK_ERR_NOMEM, "mgpu: out of host memory");
    g->max_cols = max_cols;
    const int32_t per = (max_cols + n_devices - 1) / n_devices;
    for (int d = 0; d < n_devices; d++) {
        bnk_ws *ws = nullptr;
        const int rc = bnk_ws_create(m, t, per, devices ? devices[d] : d, nullptr, &ws);
        if (rc != BNK_OK) { bnk_mgpu_destroy(g); return rc; }
        g->ws.push_back(ws);
        g->dev.push_back(ws->device);
    }
    *out = g;
    return BNK_OK;
}

extern "C" void bnk_mgpu_destroy(bnk_mgpu *g)
{
    if (!g) return;
    for (size_t d = 0; d < g->comms.size(); d++)
        if (g->comms[d]) g->nccl.CommDestroy(g->comms[d]);
    for (bnk_ws *ws : g->ws) ws_free(ws);
    delete g;
}

extern "C" int bnk_mgpu_n_devices(const bnk_mgpu *g) { return g ? (int)g->ws.size() : fail(BNK_ERR_ARG, "mgpu: null"); }

extern "C" int bnk_mgpu_shard(const bnk_mgpu *g, int32_t n_cols, int idx, int32_t *col0, int32_t *ncols)
{
    if (!g || idx < 0 || idx >= (int)g->ws.size() || n_cols < 0) return fail(BNK_ERR_ARG, "mgpu shard: bad argument");
    int32_t c0, nc;
    shard_of((int)g->ws.size(), n_cols, idx, c0, nc);
    if (col0) *col0 = c0;
    if (ncols) *ncols = nc;
    return BNK_OK;
}

extern "C" const char *bnk_mgpu_reduction(const bnk_mgpu *g) { return g ? g->reduce_how.c_str() : ""; }

extern "C" int bnk_mgpu_set_rates(bnk_mgpu *g, int32_t n_cols, const double *rates)
{
    if (!g || n_cols < 0 || n_cols > g->max_cols) return fail(BNK_ERR_ARG, "mgpu set_rates: bad argument");
    if (rates) {
        for (int32_t c = 0; c < n_cols; c++)
            if (!(rates[c] >= 0.0) || std::isinf(rates[c])) return fail(BNK_ERR_ARG, "set_rates: rate[%d] is negative or not finite", c);
        g->rates.assign(rates, rates + n_cols);
    } else g->rates.clear();
    if (g->res_cols == n_cols)  // resident shards follow
        for (size_t d = 0; d < g->ws.size(); d++) {
            int32_t c0, nc;
            shard_of((int)g->ws.size(), n_cols, (int)d, c0, nc);
            if (nc > 0) TRY(bnk_set_rates(g->ws[d], nc, rates ? rates + c0 : nullptr));
        }
    return BNK_OK;
}

extern "C" int bnk_mgpu_set_distances(bnk_mgpu *g, const double *dist)
{
    if (!g || !dist) return fail(BNK_ERR_ARG, "mgpu set_distances: bad argument");
    for (bnk_ws *ws : g->ws) TRY(bnk_set_distances(ws, dist));
    return BNK_OK;
}

// one host thread per device; each runs its column block through that device's chunk pipeline
static int mgpu_run(bnk_mgpu *g, HostCall proto)
{
    const int G = (int)g->ws.size();
    if (proto.n_cols_total < 1 || proto.n_cols_total > g->max_cols) return fail(BNK_ERR_ARG, "mgpu: n_cols %d out of range (max_cols %d)", proto.n_cols_total, g->max_cols);
    if (!g->rates.empty() && (int32_t)g->rates.size() != proto.n_cols_total)
        return fail(BNK_ERR_STATE, "mgpu: rates were set for %d columns, the call has %d", (int)g->rates.size(), proto.n_cols_total);
    proto.rates = g->rates.empty() ? nullptr : g->rates.data();
    std::vector<int> rcs(G, BNK_OK);
    std::vector<std::string> msgs(G);
    std::vector<double> sums(G, 0.0);
    std::vector<std::thread> th;
    for (int d = 0; d < G; d++) {
        th.emplace_back([&, d]() {
            HostCall hc = proto;
            shard_of(G, proto.n_cols_total, d, hc.col0, hc.ncols);
            hc.out_sum = proto.out_sum ? &sums[d] : nullptr;
            if (hc.ncols < 1) return;
            rcs[d] = ws_run_host_range(g->ws[d], hc);
            if (rcs[d] != BNK_OK) msgs[d] = bnk_last_error();
        });
    }
    for (auto &t : th) t.join();
    for (int d = 0; d < G; d++)
        if (rcs[d] != BNK_OK) return fail(rcs[d], "device %d: %s", g->dev[d], msgs[d].c_str());
    if (proto.out_sum) {
        double s = 0.0;
        for (double x : sums) s += x;
        *proto.out_sum = s;
        g->reduce_how = "host add of the per-device sums (host-buffer call)";
    }
    return BNK_OK;
}

extern "C" int bnk_mgpu_joint(bnk_mgpu *g, int32_t n_cols, const uint8_t *obs, const uint8_t *present, uint8_t *out_state)
{
    if (!g || !obs || !out_state) return fail(BNK_ERR_ARG, "mgpu joint: null argument");
    return mgpu_run(g, HostCall{0, n_cols, 0, 0, obs, present, nullptr, out_state, nullptr, 0, nullptr, nullptr, nullptr});
}

extern "C" int bnk_mgpu_marginal(bnk_mgpu *g, int32_t n_cols, const uint8_t *obs, const uint8_t *present,
                                 const int32_t *query_nodes, int32_t n_query, double *out_post, double *out_logl)
{
    if (!g || !obs) return fail(BNK_ERR_ARG, "mgpu marginal: null argument");
    return mgpu_run(g, HostCall{1, n_cols, 0, 0, obs, present, nullptr, nullptr, query_nodes, n_query, out_post, out_logl, nullptr});
}

extern "C" int bnk_mgpu_loglik(bnk_mgpu *g, int32_t n_cols, const uint8_t *obs, const uint8_t *present, double *out_logl,
                               double *out_sum)
{
    if (!g || !obs) return fail(BNK_ERR_ARG, "mgpu loglik: null argument");
    double total = 0.0;
    TRY(mgpu_run(g, HostCall{2, n_cols, 0, 0, obs, present, nullptr, nullptr, nullptr, 0, nullptr, out_logl, &total}));
    if (out_sum) *out_sum = total;
    return BNK_OK;
}

// ---- optimisation loops: the alignment stays on the devices, every evaluation is kernels + one all-reduce ----
extern "C" int bnk_mgpu_upload(bnk_mgpu *g, int32_t n_cols, const uint8_t *obs, const uint8_t *present)
{
    if (!g || !obs || n_cols < 1 || n_cols > g->max_cols) return fail(BNK_ERR_ARG, "mgpu upload: bad argument");
    if (!g->rates.empty() && (int32_t)g->rates.size() != n_cols)
        return fail(BNK_ERR_STATE, "mgpu: rates were set for %d columns, the alignment has %d", (int)g->rates.size(), n_cols);
    const int G = (int)g->ws.size();
    for (int d = 0; d < G; d++) {
        bnk_ws *ws = g->ws[d];
        int32_t c0, nc;
        shard_of(G, n_cols, d, c0, nc);
        if (nc < 1) continue;
        ENTER_DEVICE(ws->device);
        TRY(bnk_set_rates(ws, nc, g->rates.empty() ? nullptr : g->rates.data() + c0));
        const size_t bytes = (size_t)ws->n * nc;
        CUDA_TRY(ws->d_obs_in.ensure(bytes));
        CUDA_TRY(cudaMemcpy2DAsync(ws->d_obs_in.p, nc, obs + c0, n_cols, nc, ws->n, cudaMemcpyHostToDevice, ws->stream));
        if (present) {
            CUDA_TRY(ws->d_present_in.ensure(bytes));
            CUDA_TRY(cudaMemcpy2DAsync(ws->d_present_in.p, nc, present + c0, n_cols, nc, ws->n, cudaMemcpyHostToDevice, ws->stream));
        }
        CUDA_TRY(ws->d_logl.ensure(nc));
        // validate once, then let the evaluations run without a host round trip
        int32_t flag = 0;
        CUDA_TRY(cudaMemsetAsync(ws->d_flag.p, 0, sizeof(int32_t), ws->stream));
        CUDA_TRY(launch_scan_obs(ws->d_obs_in.p, ws->d_has_child.p, ws->n, nc, ws->K, ws->d_flag.p, ws->stream));
        ws->launches++;
        CUDA_TRY(cudaMemcpyAsync(&flag, ws->d_flag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ws->stream));
        CUDA_TRY(cudaStreamSynchronize(ws->stream));
        if (flag & 2) return fail(BNK_ERR_ARG, "observed states must be < %d (the model's alphabet size) or BNK_NONE (255)", ws->K);
        ws->evidence_mode = (flag & 1) ? BNK_EVIDENCE_ANYWHERE : BNK_EVIDENCE_LEAVES;
    }
    g->res_cols = n_cols;
    g->res_present = present != nullptr;
    return BNK_OK;
}

static void mgpu_init_nccl(bnk_mgpu *g)
{
    if (g->nccl_tried) return;
    g->nccl_tried = true;
    const int G = (int)g->ws.size();
    if (G < 2) { g->reduce_how = "single device: no reduction"; return; }
    if (!g->nccl.load()) { g->reduce_how = "host add of the per-device sums (libnccl.so.2 not loadable)"; return; }
    g->comms.assign(G, nullptr);
    if (g->nccl.CommInitAll(g->comms.data(), G, g->dev.data()) != ncclSuccess) {
        g->comms.clear();
        g->reduce_how = "host add of the per-device sums (ncclCommInitAll failed)";
        return;
    }
    g->reduce_how = "ncclAllReduce(sum, 1 x f64) over " + std::to_string(G) + " devices";
}

extern "C" int bnk_mgpu_loglik_resident(bnk_mgpu *g, double *out_sum)
{
    if (!g || !out_sum) return fail(BNK_ERR_ARG, "mgpu loglik: null argument");
    if (g->res_cols < 1) return fail(BNK_ERR_STATE, "mgpu loglik: call bnk_mgpu_upload first");
    mgpu_init_nccl(g);
    const int G = (int)g->ws.size();
    std::vector<int> active;
    for (int d = 0; d < G; d++) {
        bnk_ws *ws = g->ws[d];
        int32_t c0, nc;
        shard_of(G, g->res_cols, d, c0, nc);
        ENTER_DEVICE(ws->device);
        if (nc < 1) { CUDA_TRY(cudaMemsetAsync(ws->d_sum.p, 0, sizeof(double), ws->stream)); continue; }
        TRY(bnk_loglik_dev(ws, nc, ws->d_obs_in.p, g->res_present ? ws->d_present_in.p : nullptr, ws->d_logl.p, ws->d_sum.p));
        active.push_back(d);
    }
    double total = 0.0;
    if (!g->comms.empty()) {
        g->nccl.GroupStart();
        for (int d = 0; d < G; d++)
            g->nccl.AllReduce(g->ws[d]->d_sum.p, g->ws[d]->d_sum.p, 1, ncclDouble, ncclSum, g->comms[d], g->ws[d]->stream);
        if (g->nccl.GroupEnd() != ncclSuccess) return fail(BNK_ERR_CUDA, "mgpu loglik: ncclAllReduce failed");
        {
            ENTER_DEVICE(g->ws[0]->device);
            CUDA_TRY(cudaMemcpyAsync(&total, g->ws[0]->d_sum.p, sizeof(double), cudaMemcpyDeviceToHost, g->ws[0]->stream));
        }
        for (int d = 0; d < G; d++) CUDA_TRY(cudaStreamSynchronize(g->ws[d]->stream));
    } else {
        std::vector<double> part(G, 0.0);
        for (int d : active) {
            ENTER_DEVICE(g->ws[d]->device);
            CUDA_TRY(cudaMemcpyAsync(&part[d], g->ws[d]->d_sum.p, sizeof(double), cudaMemcpyDeviceToHost, g->ws[d]->stream));
        }
        for (int d = 0; d < G; d++) CUDA_TRY(cudaStreamSynchronize(g->ws[d]->stream));
        for (double p : part) total += p;
    }
    *out_sum = total;
    return BNK_OK;
}

extern "C" int64_t bnk_mgpu_launch_count(const bnk_mgpu *g)
{
    int64_t l = 0;
    if (g) for (const bnk_ws *ws : g->ws) l += bnk_ws_launch_count(ws);
    return l;
}
