// api.cu -- the C ABI of libbnkgpu.so (include/bnkgpu.h): workspace, rate categories, column
// tiling, and the orchestration of the device kernels.  No CPU fallback: every compute entry
// point returns BNK_ERR_CUDA if the device cannot run it.
#include "device_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

namespace bnk {

static thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

}  // namespace bnk

#include "workspace.cuh"

using namespace bnk;

// --------------------------------------------------------------------------------------
extern "C" const char *bnk_last_error(void) { return g_err.c_str(); }
extern "C" const char *bnk_version(void) { return "bnkgpu 0.1 (sm_100a)"; }

extern "C" int bnk_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return fail(BNK_ERR_CUDA, "no CUDA runtime / device"); }
    int ok = 0;
    for (int d = 0; d < n; d++) {
        int major = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d);
        if (major == 10) ok++;
    }
    return ok;
}

extern "C" int bnk_model_create(const char *name, const double *F, bnk_model **out) { return model_by_name(name, F, out); }
extern "C" int bnk_model_create_custom(int K, const double *F, const double *IRM, int symmetric, int normalise, bnk_model **out)
{
    return model_from_rates(K, F, IRM, symmetric != 0, normalise != 0, out);
}
extern "C" int bnk_model_create_from_eigen(int K, const double *F, const double *evec, const double *ievec,
                                           const double *eval, bnk_model **out)
{
    return model_from_eigen(K, F, evec, ievec, eval, out);
}
extern "C" int bnk_ws_dims(const bnk_ws *ws, int32_t *n_nodes, int32_t *n_internal, int32_t *n_states)
{
    if (!ws) return fail(BNK_ERR_ARG, "ws: null");
    if (n_nodes) *n_nodes = ws->n;
    if (n_internal) *n_internal = ws->n_int;
    if (n_states) *n_states = ws->K;
    return BNK_OK;
}
extern "C" int bnk_indel_base_model(const char *name, bnk_model **out) { return indel_base_model_by_name(name, out); }
extern "C" void bnk_model_destroy(bnk_model *m) { delete m; }
extern "C" int bnk_model_nstates(const bnk_model *m) { return m ? m->K : fail(BNK_ERR_ARG, "model: null"); }
extern "C" int bnk_model_get(const bnk_model *m, double *R, double *evec, double *ievec, double *eval, double *F, double *prior)
{
    if (!m) return fail(BNK_ERR_ARG, "model: null");
    const int K = m->K;
    if (R) {
        if (!m->has_R) return fail(BNK_ERR_STATE, "model: created from an eigensystem, no rate matrix kept");
        std::memcpy(R, m->R, sizeof(double) * K * K);
    }
    if (evec) std::memcpy(evec, m->evec, sizeof(double) * K * K);
    if (ievec) std::memcpy(ievec, m->ievec, sizeof(double) * K * K);
    if (eval) std::memcpy(eval, m->eval, sizeof(double) * K);
    if (F) std::memcpy(F, m->F, sizeof(double) * K);
    if (prior) std::memcpy(prior, m->prior, sizeof(double) * K);
    return BNK_OK;
}

// --------------------------------------------------------------------------------------
// workspace
// --------------------------------------------------------------------------------------
template <typename T>
static int upload(DevBuf<T> &b, const std::vector<T> &v, cudaStream_t st)
{
    CUDA_TRY(b.ensure(v.size() ? v.size() : 1));
    if (!v.empty()) CUDA_TRY(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return BNK_OK;
}

void bnk::ws_free(bnk_ws *ws)
{
    if (!ws) return;
    DeviceGuard guard(ws->device);
    if (ws->stream) cudaStreamSynchronize(ws->stream);
    for (bnk_ws *sub : ws->subs) ws_free(sub);
    if (ws->twin) ws_free(ws->twin);
    if (ws->twin_fork) cudaEventDestroy(ws->twin_fork);
    if (ws->twin_join) cudaEventDestroy(ws->twin_join);
    if (ws->copy_stream) { cudaStreamSynchronize(ws->copy_stream); cudaStreamDestroy(ws->copy_stream); }
    for (auto &e : ws->chunk_done) if (e) cudaEventDestroy(e);
    for (auto &e : ws->copy_done) if (e) cudaEventDestroy(e);
    ws->subs.clear();
    for (DevBufBase *b : ws->reg.all) b->release();
    for (auto &e : ws->ev) for (auto &x : e) if (x) cudaEventDestroy(x);
    if (ws->own_stream && ws->stream) cudaStreamDestroy(ws->stream);
    if (ws->tree) tree_release(const_cast<bnk_tree *>(ws->tree));  // the workspace held a reference
    delete ws;
}

extern "C" int bnk_ws_create(const bnk_model *m, const bnk_tree *t, int32_t max_cols, int device, void *stream, bnk_ws **out)
{
    if (!m || !t || !out || max_cols < 1) return fail(BNK_ERR_ARG, "ws: bad argument");
    if (m->K != 20 && m->K != 4) return fail(BNK_ERR_MODEL, "ws: device kernels exist for 20- and 4-state models, got %d", m->K);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(BNK_ERR_CUDA, "ws: no CUDA device visible (libbnkgpu has no CPU fallback)");
    }
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));
    ENTER_DEVICE(device);
    int major = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) return fail(BNK_ERR_CUDA, "ws: device %d is sm_%dx; libbnkgpu is built for sm_100a only", device, major);
    bnk_ws *ws = new (std::nothrow) bnk_ws();
    if (!ws) return fail(BNK_ERR_NOMEM, "ws: out of host memory");
    ws->model = *m;
    ws->tree = t;
    tree_retain(const_cast<bnk_tree *>(t));
    ws->K = m->K; ws->n = t->n; ws->n_int = t->n_int; ws->max_cols = max_cols; ws->device = device;
    cudaDeviceGetAttribute(&ws->sm_count, cudaDevAttrMultiProcessorCount, device);
    ws->smem_optin = walk_max_smem_optin(device);
    if (stream) ws->stream = (cudaStream_t)stream;
    else {
        cudaError_t e = cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { ws->stream = nullptr; ws_free(ws); return fail(BNK_ERR_CUDA, "ws: cudaStreamCreate: %s", cudaGetErrorString(e)); }
        ws->own_stream = true;
    }
    for (auto &e : ws->ev) for (auto &x : e) cudaEventCreate(&x);
    if (const char *env = std::getenv("BNK_NO_CHERRY")) ws->no_cherry = std::atoi(env) != 0;
    if (const char *env = std::getenv("BNK_NO_FUSE2")) ws->no_fuse2 = std::atoi(env) != 0;
    if (const char *env = std::getenv("BNK_CHERRY_DIRECT")) ws->no_cherry_direct = std::atoi(env) == 0;
    if (const char *env = std::getenv("BNK_WALKER"))
        ws->walker_mode = std::strcmp(env, "lean") == 0 ? 1 : (std::strcmp(env, "scalar") == 0 ? 2 : (std::strcmp(env, "warp") == 0 ? 3 : (std::strcmp(env, "general") == 0 ? 5 : 0)));
    if (const char *env = std::getenv("BNK_MMA_W")) ws->mw = -std::atoi(env);  // negative: forced
    if (const char *env = std::getenv("BNK_WARP_W")) ws->cfg_xw = std::atoi(env);
    if (const char *env = std::getenv("BNK_WARP_NS")) ws->cfg_xns = std::atoi(env);
    if (const char *env = std::getenv("BNK_WARP_C")) ws->cfg_xc = std::atoi(env);
    if (const char *env = std::getenv("BNK_WIDE_TPC")) ws->wide_tpc_max = std::max(1, std::atoi(env));  // tuning knob
    if (const char *env = std::getenv("BNK_TABLE_BUDGET_MB"))  // bytes of P(t) tables held at once (tests force several chunks)
        ws->table_budget = (size_t)std::max(1, std::atoi(env)) << 20;

    const int K = ws->K;
    std::vector<double> mh(2 * K * K + 3 * K, 0.0);
    std::memcpy(&mh[0], m->evec, sizeof(double) * K * K);
    std::memcpy(&mh[K * K], m->ievec, sizeof(double) * K * K);
    std::memcpy(&mh[2 * K * K], m->eval, sizeof(double) * K);
    std::memcpy(&mh[2 * K * K + K], m->prior, sizeof(double) * K);
    std::vector<int32_t> leaf_nodes, leaf_parent;
    std::vector<uint8_t> has_child(t->n, 0);
    for (int32_t v = 0; v < t->n; v++) {
        if (t->ent_of_node[v] < 0) { leaf_nodes.push_back(v); leaf_parent.push_back(t->parent[v]); }
        else has_child[v] = 1;
    }
    ws->n_leaves = (int)leaf_nodes.size();
    ws->dist = t->dist;
    int rc = BNK_OK;
    auto step = [&](int r) { if (rc == BNK_OK) rc = r; };
    step(upload(ws->d_model, mh, ws->stream));
    step(upload(ws->d_up, t->full.up, ws->stream));
    step(upload(ws->d_down, t->full.down, ws->stream));
    step(upload(ws->d_up_child, t->full.up_child, ws->stream));
    step(upload(ws->d_down_child, t->full.down_child, ws->stream));
    step(upload(ws->d_up_n, t->narrow.up, ws->stream));
    step(upload(ws->d_down_n, t->narrow.down, ws->stream));
    step(upload(ws->d_up_child_n, t->narrow.up_child, ws->stream));
    step(upload(ws->d_down_child_n, t->narrow.down_child, ws->stream));
    std::vector<WideNode> wide_all;
    ws->level_off.assign(1, 0);
    for (const auto &lv : t->levels) {
        wide_all.insert(wide_all.end(), lv.begin(), lv.end());
        ws->level_off.push_back((int32_t)wide_all.size());
    }
    step(upload(ws->d_wide, wide_all, ws->stream));
    if (t->wide_h >= 2) {  // the fused down-sweep: every wide node above height 1 evaluates its height-1 children itself
        const auto &l1 = t->levels[0];
        std::vector<int32_t> row_of_node(t->n, -1);
        for (size_t i = 0; i < l1.size(); i++) row_of_node[l1[i].node] = (int32_t)i;
        std::vector<uint8_t> taken(l1.size(), 0);
        std::vector<Wide2Node> f2;
        ws->fuse_off.assign(t->wide_h + 1, 0);
        for (int h = 1; h < t->wide_h; h++) {
            ws->fuse_off[h] = (int32_t)f2.size();
            for (const WideNode &v : t->levels[h]) {
                Wide2Node f;
                f.v = v;
                for (int e = 0; e < 2; e++) {
                    std::memset(&f.c[e], 0, sizeof(WideNode));
                    f.c[e].node = -1;
                    const int32_t u = e < v.n_child ? v.c_node[e] : -1;
                    f.crow[e] = -1;
                    if (u >= 0 && row_of_node[u] >= 0) { f.c[e] = l1[row_of_node[u]]; f.crow[e] = row_of_node[u]; taken[row_of_node[u]] = 1; }
                }
                f2.push_back(f);
            }
        }
        ws->fuse_off[t->wide_h] = (int32_t)f2.size();
        std::vector<WideNode> rest;  // height-1 nodes under a walked (narrow) parent keep the plain level kernel
        std::vector<int32_t> rest_rows, cherry_row(std::max(t->n_int, 1), -1);
        for (size_t i = 0; i < l1.size(); i++) {
            if (!taken[i]) { rest.push_back(l1[i]); rest_rows.push_back((int32_t)i); }
            else cherry_row[l1[i].ent] = (int32_t)i;  // cherry-direct: its message stays in the record tables
        }
        ws->n_fuse2 = (int)f2.size(); ws->n_wide_rest = (int)rest.size();
        step(upload(ws->d_fuse2, f2, ws->stream));
        step(upload(ws->d_wide_rest, rest, ws->stream));
        step(upload(ws->d_rest_rows, rest_rows, ws->stream));
        step(upload(ws->d_cherry_row, cherry_row, ws->stream));
    }
    step(upload(ws->d_parent, t->parent, ws->stream));
    step(upload(ws->d_node_of_ent, t->node_of_ent, ws->stream));
    step(upload(ws->d_leaf_nodes, leaf_nodes, ws->stream));
    step(upload(ws->d_leaf_parent, leaf_parent, ws->stream));
    step(upload(ws->d_has_child, has_child, ws->stream));
    step(upload(ws->d_dist, ws->dist, ws->stream));
    if (rc == BNK_OK && ws->d_flag.ensure(1) != cudaSuccess) rc = fail(BNK_ERR_NOMEM, "ws: device allocation failed");
    if (rc == BNK_OK && ws->d_sum.ensure(1) != cudaSuccess) rc = fail(BNK_ERR_NOMEM, "ws: device allocation failed");
    if (rc == BNK_OK) {
        launch_log_prior(ws->md().prior, const_cast<double *>(ws->md().logprior), K, ws->stream);
        ws->launches++;
        cudaError_t e = cudaStreamSynchronize(ws->stream);
        if (e != cudaSuccess) rc = fail(BNK_ERR_CUDA, "ws: initialisation failed: %s", cudaGetErrorString(e));
    }
    if (rc != BNK_OK) { ws_free(ws); return rc; }
    *out = ws;
    return BNK_OK;
}

extern "C" void bnk_ws_destroy(bnk_ws *ws) { ws_free(ws); }

extern "C" int bnk_ws_sync(bnk_ws *ws)
{
    if (!ws) return fail(BNK_ERR_ARG, "ws: null");
    CUDA_TRY(cudaStreamSynchronize(ws->stream));
    return BNK_OK;
}

extern "C" int bnk_ws_configure(bnk_ws *ws, int tile_cols, int cols_per_thread)
{
    if (!ws) return fail(BNK_ERR_ARG, "ws: null");
    ws->force_generic = cols_per_thread < 0;  // negative: same geometry, generic kernels only (tests)
    if (cols_per_thread < 0) cols_per_thread = -cols_per_thread;
    if (cols_per_thread != 0 && cols_per_thread != 1 && cols_per_thread != 2 && cols_per_thread != 4)
        return fail(BNK_ERR_ARG, "ws: cols_per_thread must be 0 (auto), 1, 2 or 4");
    if (ws->K == 4 && cols_per_thread > 1) return fail(BNK_ERR_ARG, "ws: 4-state kernels are built with 1 column per thread");
    ws->no_wide = tile_cols < 0;  // negative: same geometry, walker kernels only (tests, profiling)
    if (tile_cols < 0) tile_cols = -tile_cols;
    ws->cfg_tile_cols = tile_cols;
    ws->cfg_cpt = cols_per_thread;
    ws->retile = true;  // tiles are rebuilt, for the rates already set, by the next call that needs them
    return BNK_OK;
}

extern "C" int bnk_ws_set_pipeline(bnk_ws *ws, int32_t chunk_cols)
{
    if (!ws) return fail(BNK_ERR_ARG, "ws: null");
    ws->pipeline_cols = chunk_cols;
    return BNK_OK;
}

extern "C" int bnk_ws_set_evidence_mode(bnk_ws *ws, int mode)
{
    if (!ws || mode < BNK_EVIDENCE_AUTO || mode > BNK_EVIDENCE_ANYWHERE) return fail(BNK_ERR_ARG, "evidence mode: bad argument");
    ws->evidence_mode = mode;
    for (bnk_ws *sub : ws->subs) sub->evidence_mode = mode;
    return BNK_OK;
}

extern "C" int64_t bnk_ws_launch_count(const bnk_ws *ws)
{
    if (!ws) return 0;
    int64_t l = ws->launches;
    for (const bnk_ws *sub : ws->subs) l += sub->launches;
    if (ws->twin) l += ws->twin->launches;
    return l;
}
extern "C" int64_t bnk_ws_device_bytes(const bnk_ws *ws)
{
    if (!ws) return 0;
    int64_t b = ws->reg.bytes;
    for (const bnk_ws *sub : ws->subs) b += sub->reg.bytes;
    if (ws->twin) b += ws->twin->reg.bytes;
    return b;
}

extern "C" int bnk_ws_phase_ms(bnk_ws *ws, int phase, float *ms)
{
    if (!ws || !ms || phase < 0 || phase > 4) return fail(BNK_ERR_ARG, "phase_ms: bad argument");
    if (ws->twin && ws->twin_phases && (phase == 1 || phase == 2)) ws = ws->twin;  // the joint pass of bnk_joint_marginal_dev
    if (!ws->ev_valid[phase]) { *ms = 0.f; return BNK_OK; }
    CUDA_TRY(cudaEventElapsedTime(ms, ws->ev[phase][0], ws->ev[phase][1]));
    return BNK_OK;
}

// --------------------------------------------------------------------------------------
// rates -> categories -> sorted columns -> tiles -> chunks
// --------------------------------------------------------------------------------------
static const int MAXCAT_TILE = 4;

extern "C" int bnk_set_rates(bnk_ws *ws, int32_t n_cols, const double *rates)
{
    if (!ws || n_cols < 0) return fail(BNK_ERR_ARG, "set_rates: bad argument");
    if (n_cols > ws->max_cols) return fail(BNK_ERR_ARG, "set_rates: n_cols %d exceeds the workspace's max_cols %d", n_cols, ws->max_cols);
    if (rates)
        for (int32_t c = 0; c < n_cols; c++)
            if (!(rates[c] >= 0.0) || std::isinf(rates[c])) return fail(BNK_ERR_ARG, "set_rates: rate[%d] is negative or not finite", c);
    ENTER_DEVICE(ws->device);
    const int K = ws->K;
    // same assignment as last time (an optimisation loop re-evaluating, a second pass over the same
    // alignment): keep categories, tiles and their device copies; only the tables are marked stale
    if (!ws->retile && ws->ncols == n_cols && ws->rates_null == (rates == nullptr) &&
        (rates == nullptr || (ws->last_rates.size() == (size_t)n_cols &&
                              std::memcmp(ws->last_rates.data(), rates, sizeof(double) * n_cols) == 0)))
        return BNK_OK;  // the tables depend on (rates, distances) only: still valid (bnk_set_distances marks them stale)
    std::vector<double> kept;  // rates may alias last_rates (re-tiling after bnk_ws_configure)
    if (rates && !ws->last_rates.empty() && rates == ws->last_rates.data()) { kept = ws->last_rates; rates = kept.data(); }
    ws->retile = false;
    // distinct rates, ascending; columns sorted by (rate, index)
    ws->orig_col.resize(n_cols);
    for (int32_t c = 0; c < n_cols; c++) ws->orig_col[c] = c;
    ws->cat_rate.clear();
    ws->col_cat_global.assign(n_cols, 0);
    ws->cat_of_col.assign(n_cols, 0);
    ws->identity_perm = true;
    if (rates && n_cols > 0) {
        std::stable_sort(ws->orig_col.begin(), ws->orig_col.end(), [&](int32_t a, int32_t b) { return rates[a] < rates[b]; });
        for (int32_t c = 0; c < n_cols; c++) {
            const double r = rates[ws->orig_col[c]];
            if (ws->cat_rate.empty() || r != ws->cat_rate.back()) ws->cat_rate.push_back(r);
            ws->col_cat_global[c] = (int32_t)ws->cat_rate.size() - 1;
            ws->cat_of_col[ws->orig_col[c]] = ws->col_cat_global[c];
            if (ws->orig_col[c] != c) ws->identity_perm = false;
        }
    } else {
        ws->cat_rate.push_back(1.0);  // PhyloBN.DEFAULT_RATE (Prediction.java:615-618)
    }
    const int ncat = (int)ws->cat_rate.size();
    // thread / tile geometry
    // defaults from the r1 sweep on C3 (profiles/): 2 columns per thread, ~2.8 CTAs per SM
    int cpt = ws->cfg_cpt ? ws->cfg_cpt : (K == 20 ? 2 : 1);
    int tc = ws->cfg_tile_cols;
    const int max_threads = cpt == 1 ? 512 : (cpt == 2 ? 384 : 256);  // the kernels' __launch_bounds__
    const int tc_cap = (max_threads / K) * cpt;
    if (tc <= 0) {
        const int target_tiles = ws->sm_count * 3 - ws->sm_count / 6;
        tc = (n_cols + target_tiles - 1) / target_tiles;
        if (tc < 4) tc = 4;
    }
    if (tc > tc_cap) tc = tc_cap;
    if (tc < 1) tc = 1;
    ws->tile_cols = tc; ws->cpt = cpt;
    ws->ncg = (tc + cpt - 1) / cpt;
    ws->threads = ((ws->ncg * K + 31) / 32) * 32;
    // tiles: big categories are split evenly; small ones are packed together (<= MAXCAT_TILE per tile)
    ws->tiles.clear();
    std::vector<int> cat_begin(ncat + 1, 0);
    for (int32_t c = 0; c < n_cols; c++) cat_begin[ws->col_cat_global[c] + 1]++;
    for (int k = 0; k < ncat; k++) cat_begin[k + 1] += cat_begin[k];
    // chunks of categories bounded by the table budget
    const size_t per_cat = (size_t)ws->n * K * K * sizeof(double) * 2;
    int cats_per_chunk = (int)std::max<size_t>(1, ws->table_budget / std::max<size_t>(per_cat, 1));
    if (cats_per_chunk > 60000) cats_per_chunk = 60000;
    ws->chunks.clear(); ws->wtiles.clear(); ws->wtile_group.clear(); ws->chunk_wtiles.clear(); ws->chunk_cols.clear();
    ws->cgroups.clear(); ws->chunk_cgroups.clear(); ws->chunk_crecs.clear(); ws->ctiles.clear(); ws->chunk_ctiles.clear();
    ws->xtiles.clear(); ws->chunk_xtiles.clear();
    {   // warp-tiled walker geometry: W consumer warps of 6 x C columns per CTA.  Every CTA of the grid should be
        // resident at once (a walk is one long dependent chain per column), so prefer the tile width that fills
        // the SMs most evenly -- cost = rounds of CTAs per SM x columns per CTA -- and, for a given width, one
        // column per lane (more warps to hide the shared-memory latency: C3b 35.5 vs 37.5 ms per step)
        int best_w = 2, best_c = 1;
        double best_cost = -1.0;
        for (int c = 1; c <= 2; c++) {
            if (ws->cfg_xc >= 1 && ws->cfg_xc <= 3 && c != ws->cfg_xc) continue;
            if (!(ws->cfg_xc >= 1 && ws->cfg_xc <= 3) && c != 1) continue;  // one column per lane unless BNK_WARP_C asks (r2 A/B on C3b: C = 2 loses, 31.4 vs 30.2 ms; the masked walker needs C = 1)
            const int cw = warp_walk_cols_per_warp(c);
            for (int w = 2; w <= 7 && w * cw <= 64; w++) {
                if (c == 2 && w > 4) continue;
                int64_t nt = 0;
                for (int k = 0; k < ncat; k++) nt += (cat_begin[k + 1] - cat_begin[k] + w * cw - 1) / (w * cw);
                const int64_t rounds = (nt + ws->sm_count - 1) / ws->sm_count;
                // at equal cost fewer, wider CTAs win (one resident CTA per SM: no second producer, no imbalance)
                const double cost = (double)rounds * w * c * (c == 1 ? 1.0 : 1.06) * (1.0 + 0.05 * (double)(rounds - 1));
                if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_w = w; best_c = c; }
            }
        }
        ws->xc = (ws->cfg_xc >= 1 && ws->cfg_xc <= 3) ? ws->cfg_xc : best_c;
        const int cw = warp_walk_cols_per_warp(ws->xc);
        if (ws->cfg_xc == 3) best_w = 2;
        ws->xw = (ws->cfg_xw >= 1 && ws->cfg_xw <= 7 && ws->cfg_xw * cw <= 64) ? ws->cfg_xw : best_w;
    }
    {   // DMMA variants: 8 columns per warp, same cost model
        int best_w = 2; double best_cost = -1.0;
        const int cw = mma_walk_cols_per_warp();
        for (int w = 2; w <= 7 && w * cw <= 64; w++) {
            int64_t nt = 0;
            for (int k = 0; k < ncat; k++) nt += (cat_begin[k + 1] - cat_begin[k] + w * cw - 1) / (w * cw);
            const int64_t rounds = (nt + ws->sm_count - 1) / ws->sm_count;
            const double cost = (double)rounds * w * (1.0 + 0.05 * (double)(rounds - 1));
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_w = w; }
        }
        if (ws->mw < 0 && -ws->mw <= 7) ws->mw = ws->mw;  // forced by BNK_MMA_W (kept negative as the marker)
        else ws->mw = best_w;
    }
    const int wt = wide_tile_cols(), cg = cherry_group_cols();
    ws->mtiles.clear(); ws->chunk_mtiles.clear();
    for (int lo = 0; lo < ncat; lo += cats_per_chunk) {
        const int hi = std::min(ncat, lo + cats_per_chunk);
        Chunk ck{lo, hi, (int)ws->tiles.size(), 0, 1};
        const int w0 = (int)ws->wtiles.size();
        for (int k = lo; k < hi; k++) {
            // Tiles of the level kernels: one category each, <= wt columns, and every tile but the first of its
            // category starts on a multiple of 32 columns of the sorted space -- so each warp's row segment (32
            // columns x 8 bytes) is one aligned 256-byte run: 8 sectors in 2 lines.  An even split (r1) left all
            // tiles of a category as misaligned as its first column: 9 sectors in 3 lines per warp and row.
            const int c_lo = cat_begin[k], c_hi = cat_begin[k + 1];
            int a0 = c_lo;
            while (a0 < c_hi) {
                int a1 = ((a0 + wt) / 32) * 32;          // the last multiple of 32 within reach
                if (a1 <= a0 || a0 % 32 == 0) a1 = a0 + wt;
                if (a1 > c_hi) a1 = c_hi;
                ws->wtiles.push_back(TileDesc{a0, a1 - a0, k - lo, 1});
                a0 = a1;
            }
        }
        const int x0 = (int)ws->xtiles.size();
        {
            const int xt = ws->xw * warp_walk_cols_per_warp(ws->xc);
            for (int k = lo; k < hi; k++) {
                const int nc = cat_begin[k + 1] - cat_begin[k];
                const int nt = (nc + xt - 1) / xt;
                for (int q = 0; q < nt; q++) {
                    const int a0 = cat_begin[k] + (int)((int64_t)nc * q / nt), a1 = cat_begin[k] + (int)((int64_t)nc * (q + 1) / nt);
                    ws->xtiles.push_back(TileDesc{a0, a1 - a0, k - lo, 1});
                }
            }
        }
        const int m0 = (int)ws->mtiles.size();
        {
            const int mt = std::abs(ws->mw) * mma_walk_cols_per_warp();
            for (int k = lo; k < hi; k++) {
                const int nc = cat_begin[k + 1] - cat_begin[k];
                const int nt = (nc + mt - 1) / mt;
                for (int q = 0; q < nt; q++) {
                    const int a0 = cat_begin[k] + (int)((int64_t)nc * q / nt), a1 = cat_begin[k] + (int)((int64_t)nc * (q + 1) / nt);
                    ws->mtiles.push_back(TileDesc{a0, a1 - a0, k - lo, 1});
                }
            }
        }
        const int g0 = (int)ws->cgroups.size();
        const int ct0 = (int)ws->ctiles.size();
        int crecs = 0;  // TileDesc::ncat of a cherry group = offset of its result records inside a node's region
        for (int k = lo; k < hi; k++) {
            const int nc = cat_begin[k + 1] - cat_begin[k];
            const int nt = (nc + cg - 1) / cg;
            for (int q = 0; q < nt; q++) {
                const int a0 = cat_begin[k] + (int)((int64_t)nc * q / nt), a1 = cat_begin[k] + (int)((int64_t)nc * (q + 1) / nt);
                for (int c = a0 & ~15; c < a1; c += wt) ws->ctiles.push_back(TileDesc{c, wt, (int)ws->cgroups.size() - g0, 0});
                ws->cgroups.push_back(TileDesc{a0, a1 - a0, k - lo, crecs});
                crecs += cherry_group_capacity(K, a1 - a0);  // distinct (state, state) pairs of two children
            }
        }
        // first cherry group (chunk-local index) overlapping every level tile of this chunk
        ws->wtile_group.resize(ws->wtiles.size(), 0);
        for (int ti = w0, gi = g0; ti < (int)ws->wtiles.size(); ti++) {
            while (gi + 1 < (int)ws->cgroups.size() && ws->wtiles[ti].col0 >= ws->cgroups[gi].col0 + ws->cgroups[gi].ncols) gi++;
            ws->wtile_group[ti] = gi - g0;
        }
        TileDesc cur{0, 0, 0, 0};
        auto close = [&]() { if (cur.ncols > 0) { ws->tiles.push_back(cur); ck.ncat_max = std::max(ck.ncat_max, cur.ncat); } cur = TileDesc{0, 0, 0, 0}; };
        for (int k = lo; k < hi; k++) {
            const int nc = cat_begin[k + 1] - cat_begin[k];
            if (nc == 0) continue;
            if (nc * 2 >= tc) {
                close();
                const int nt = (nc + tc - 1) / tc;
                for (int q = 0; q < nt; q++) {
                    const int a0 = cat_begin[k] + (int)((int64_t)nc * q / nt), a1 = cat_begin[k] + (int)((int64_t)nc * (q + 1) / nt);
                    ws->tiles.push_back(TileDesc{a0, a1 - a0, k - lo, 1});
                }
            } else {
                if (cur.ncols > 0 && (cur.ncols + nc > tc || cur.ncat >= MAXCAT_TILE || cur.col0 + cur.ncols != cat_begin[k])) close();
                if (cur.ncols == 0) { cur.col0 = cat_begin[k]; cur.cat0 = k - lo; }
                cur.ncols += nc;
                cur.ncat = (k - lo) - cur.cat0 + 1;
            }
        }
        close();
        ck.tile_hi = (int)ws->tiles.size();
        if (ck.tile_hi > ck.tile_lo) {
            ws->chunks.push_back(ck);
            ws->chunk_wtiles.push_back({w0, (int)ws->wtiles.size()});
            ws->chunk_cgroups.push_back({g0, (int)ws->cgroups.size()});
            ws->chunk_crecs.push_back(crecs);
            ws->chunk_ctiles.push_back({ct0, (int)ws->ctiles.size()});
            ws->chunk_xtiles.push_back({x0, (int)ws->xtiles.size()});
            ws->chunk_mtiles.push_back({m0, (int)ws->mtiles.size()});
            ws->chunk_cols.push_back({cat_begin[lo], cat_begin[hi]});
        }
    }
    // chunk-local category per sorted column
    std::vector<int32_t> col_cat_local(n_cols);
    for (const Chunk &ck : ws->chunks)
        for (int tI = ck.tile_lo; tI < ck.tile_hi; tI++) {
            const TileDesc &td = ws->tiles[tI];
            for (int c = td.col0; c < td.col0 + td.ncols; c++) col_cat_local[c] = ws->col_cat_global[c] - ck.cat_lo;
        }
    TRY(upload(ws->d_cat_rate, ws->cat_rate, ws->stream));
    TRY(upload(ws->d_orig_col, ws->orig_col, ws->stream));
    TRY(upload(ws->d_col_cat, col_cat_local, ws->stream));
    TRY(upload(ws->d_cat_of_col, ws->cat_of_col, ws->stream));
    TRY(upload(ws->d_tiles, ws->tiles, ws->stream));
    TRY(upload(ws->d_wtiles, ws->wtiles, ws->stream));
    TRY(upload(ws->d_wtile_group, ws->wtile_group, ws->stream));
    TRY(upload(ws->d_cgroups, ws->cgroups, ws->stream));
    TRY(upload(ws->d_ctiles, ws->ctiles, ws->stream));
    TRY(upload(ws->d_xtiles, ws->xtiles, ws->stream));
    TRY(upload(ws->d_mtiles, ws->mtiles, ws->stream));
    CUDA_TRY(cudaStreamSynchronize(ws->stream));  // host vectors above are about to go out of scope / be reused
    ws->ncols = n_cols;
    ws->rates_null = rates == nullptr;
    if (rates) ws->last_rates.assign(rates, rates + n_cols); else ws->last_rates.clear();
    ws->tables_log_valid = ws->tables_lin_valid = false;
    return BNK_OK;
}

extern "C" int bnk_set_distances(bnk_ws *ws, const double *dist)
{
    if (!ws || !dist) return fail(BNK_ERR_ARG, "set_distances: bad argument");
    for (int32_t i = 0; i < ws->n; i++)
        if (ws->tree->parent[i] >= 0 && !(dist[i] >= 0.0)) return fail(BNK_ERR_TREE, "set_distances: distance[%d] is negative or NaN", i);
    ENTER_DEVICE(ws->device);
    ws->dist.assign(dist, dist + ws->n);
    ws->dist_version++;
    CUDA_TRY(cudaMemcpyAsync(ws->d_dist.p, ws->dist.data(), sizeof(double) * ws->n, cudaMemcpyHostToDevice, ws->stream));
    CUDA_TRY(cudaStreamSynchronize(ws->stream));
    ws->tables_log_valid = ws->tables_lin_valid = false;
    return BNK_OK;
}

extern "C" int bnk_probs(bnk_ws *ws, int32_t n_t, const double *t, double *P, double *logP)
{
    if (!ws || !t || n_t < 1) return fail(BNK_ERR_ARG, "probs: bad argument");
    ENTER_DEVICE(ws->device);
    const int KK = ws->K * ws->K;
    DevBuf<double> dt, dP, dL, dr;
    int rc = BNK_OK;
    auto body = [&]() -> int {
        CUDA_TRY(dt.ensure(n_t));
        CUDA_TRY(dr.ensure(1));
        if (P) CUDA_TRY(dP.ensure((size_t)n_t * KK));
        if (logP) CUDA_TRY(dL.ensure((size_t)n_t * KK));
        const double one = 1.0;
        CUDA_TRY(cudaMemcpyAsync(dt.p, t, sizeof(double) * n_t, cudaMemcpyHostToDevice, ws->stream));
        CUDA_TRY(cudaMemcpyAsync(dr.p, &one, sizeof(double), cudaMemcpyHostToDevice, ws->stream));
        TableSet ts{dP.p, nullptr, dL.p, nullptr};
        launch_build_tables(ws->K, ws->md(), dt.p, nullptr, dr.p, n_t, 1, ts, ws->stream);
        ws->launches++;
        CUDA_TRY(cudaGetLastError());
        if (P) CUDA_TRY(cudaMemcpyAsync(P, dP.p, sizeof(double) * n_t * KK, cudaMemcpyDeviceToHost, ws->stream));
        if (logP) CUDA_TRY(cudaMemcpyAsync(logP, dL.p, sizeof(double) * n_t * KK, cudaMemcpyDeviceToHost, ws->stream));
        CUDA_TRY(cudaStreamSynchronize(ws->stream));
        return BNK_OK;
    };
    rc = body();
    dt.release(); dP.release(); dL.release(); dr.release();
    return rc;
}

// --------------------------------------------------------------------------------------
// run helpers
// --------------------------------------------------------------------------------------
static int ensure_rates(bnk_ws *ws, int32_t n_cols)
{
    if (n_cols < 1) return fail(BNK_ERR_ARG, "n_cols must be >= 1");
    if (ws->ncols < 0) return bnk_set_rates(ws, n_cols, nullptr);
    if (ws->retile) TRY(bnk_set_rates(ws, ws->ncols, ws->rates_null ? nullptr : ws->last_rates.data()));
    if (ws->ncols != n_cols)
        return fail(BNK_ERR_STATE, "n_cols = %d but rates were set for %d columns (call bnk_set_rates first)", n_cols, ws->ncols);
    return BNK_OK;
}

// Decides lean vs general kernels.  BNK_EVIDENCE_AUTO scans the observation matrix on the device (one launch):
// evidence on a node with children selects the general kernels, a byte that is neither BNK_NONE nor a state
// index is an error -- that costs one 4-byte read-back and a stream synchronisation per call.  The other two
// modes take the caller's word (bnk_ws_set_evidence_mode), launch nothing and never synchronise.
static int prepare_inputs(bnk_ws *ws, const uint8_t *d_obs, const uint8_t *d_present, bool &general,
                          const uint8_t *&obs_s, const uint8_t *&present_s)
{
    const int nc = ws->ncols;
    general = d_present != nullptr || ws->evidence_mode == BNK_EVIDENCE_ANYWHERE;
    // a presence mask with evidence on childless nodes only (what GRASP produces: per-column forests over extants'
    // residues) keeps a fast walker for deep trees: the masked warp-tiled walker of sweeps_warp.cu
    ws->masked_only = d_present != nullptr && ws->evidence_mode != BNK_EVIDENCE_ANYWHERE;
    if (ws->evidence_mode == BNK_EVIDENCE_AUTO) {
        int32_t flag = 0;
        CUDA_TRY(cudaMemsetAsync(ws->d_flag.p, 0, sizeof(int32_t), ws->stream));
        CUDA_TRY(launch_scan_obs(d_obs, ws->d_has_child.p, ws->n, nc, ws->K, ws->d_flag.p, ws->stream));
        ws->launches++;
        CUDA_TRY(cudaMemcpyAsync(&flag, ws->d_flag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ws->stream));
        CUDA_TRY(cudaStreamSynchronize(ws->stream));
        if (flag & 2)
            return fail(BNK_ERR_ARG, "observed states must be < %d (the model's alphabet size) or BNK_NONE (255)", ws->K);
        if (flag & 1) { general = true; ws->masked_only = false; }
    }
    // inputs are consumed in the caller's column order (the kernels index them through orig_col)
    obs_s = d_obs; present_s = d_present;
    return BNK_OK;
}

static int build_tables(bnk_ws *ws, const Chunk &ck, bool log_space)
{
    const int K = ws->K, KK = K * K, ncat = ck.cat_hi - ck.cat_lo;
    const size_t elems = (size_t)ncat * ws->n * KK;
    const bool single = ws->chunks.size() == 1;
    bool &valid = log_space ? ws->tables_log_valid : ws->tables_lin_valid;
    if (single && valid) return BNK_OK;
    TableSet ts{nullptr, nullptr, nullptr, nullptr};
    if (log_space) {
        CUDA_TRY(ws->d_L.ensure(elems));
        CUDA_TRY(ws->d_LT.ensure(elems));
        ts.L = ws->d_L.p; ts.LT = ws->d_LT.p;
    } else {
        CUDA_TRY(ws->d_P.ensure(elems));
        CUDA_TRY(ws->d_PT.ensure(elems));
        ts.P = ws->d_P.p; ts.PT = ws->d_PT.p;
    }
    // a workspace that has already served the other kind of pass gets both table sets from one launch
    // (the P(t) entries are computed either way; a joint + marginal step then builds tables once)
    bool &other_valid = log_space ? ws->tables_lin_valid : ws->tables_log_valid;
    if (single && !other_valid) {
        if (log_space && ws->d_P.cap >= elems && ws->d_PT.cap >= elems) { ts.P = ws->d_P.p; ts.PT = ws->d_PT.p; other_valid = true; }
        if (!log_space && ws->d_L.cap >= elems && ws->d_LT.cap >= elems) { ts.L = ws->d_L.p; ts.LT = ws->d_LT.p; other_valid = true; }
    }
    CUDA_TRY(cudaEventRecord(ws->ev[0][0], ws->stream));
    launch_build_tables(K, ws->md(), ws->d_dist.p, ws->d_parent.p, ws->d_cat_rate.p + ck.cat_lo, ws->n, ncat, ts, ws->stream);
    ws->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ws->ev[0][1], ws->stream));
    ws->ev_valid[0] = true;
    valid = single;
    return BNK_OK;
}

struct WalkSetup {
    WalkArgs a;
    WalkLaunch wl;
    bool fast = false;    // lean kernels of sweeps_fast.cu apply
    bool hybrid = false;  // wide height levels run as level launches (wide.cu); the walker takes the narrow top
    bool wmask = false;   // the masked warp-tiled walker takes the walk (s.warp is set too; s.fast stays false)
    bool warp = false;    // fast, and the warp-tiled kernels of sweeps_warp.cu take the walk
    int xw = 0, xc = 0, xns = 0, n_xtiles = 0;
    bool mma = false;     // warp, and the sum-product sweeps run the DMMA variants
    int mw = 0, mns = 0, n_mtiles = 0;
    const TileDesc *mtiles = nullptr;
    int tb_slots = 0;
};

// Decides the schedule for one chunk: lean or general walker kernels, and whether the wide
// height levels run as level launches (hybrid) with the walker restricted to the narrow top.
static int setup_walk(bnk_ws *ws, const Chunk &ck, bool general, bool want_down, const uint8_t *obs_s,
                      const uint8_t *present_s, WalkSetup &s)
{
    const size_t ci = (size_t)(&ck - ws->chunks.data());
    const int K = ws->K;
    std::memset(&s.a, 0, sizeof(s.a));
    const bnk_tree *t = ws->tree;
    s.a.tiles = ws->d_tiles.p + ck.tile_lo;
    s.a.ncg = ws->ncg;
    s.a.ncat_max = ck.ncat_max;
    s.a.obs = obs_s; s.a.present = present_s;
    s.a.ncols = ws->ncols; s.a.n_nodes = ws->n;
    s.a.orig_col = ws->d_orig_col.p; s.a.col_cat = ws->d_col_cat.p;
    s.wl.K = K; s.wl.cpt = ws->cpt; s.wl.threads = ws->threads; s.wl.n_tiles = ck.tile_hi - ck.tile_lo;
    s.wl.general = general; s.wl.stream = ws->stream;
    auto use_program = [&](bool narrow) {
        const Program &pg = narrow ? t->narrow : t->full;
        s.a.up = narrow ? ws->d_up_n.p : ws->d_up.p;
        s.a.down = narrow ? ws->d_down_n.p : ws->d_down.p;
        s.a.up_child = narrow ? ws->d_up_child_n.p : ws->d_up_child.p;
        s.a.down_child = narrow ? ws->d_down_child_n.p : ws->d_down_child.p;
        s.a.n_steps = (int32_t)pg.up.size();
        s.tb_slots = pg.tb_slots;
        return std::max(pg.up_slots, want_down ? pg.down_slots : 0);
    };
    // lean kernels: nothing masked, evidence on childless nodes only, binary (or unary) nodes,
    // one category per tile, whole stack in shared memory
    s.fast = false; s.hybrid = false;
    s.warp = false;
    s.mma = false;
    if (general && ws->masked_only && !ws->force_generic && t->max_children <= 2 && K == 20 && ws->walker_mode != 5 &&
        ws->walker_mode != 1 && ws->xc == 1) {
        // masked warp-tiled walker (r2): per-column forests in the fast deep-tree walker.  A pre-pass turns the presence
        // mask into one case byte per (walker node, column); the wide levels keep the general level kernels
        const bool hybrid = t->wide_h > 0 && !ws->no_wide;
        const int slots = use_program(hybrid);
        int ns = ws->cfg_xns >= 2 && ws->cfg_xns <= 8 ? ws->cfg_xns : 8;
        if (!(ws->cfg_xns >= 2 && ws->cfg_xns <= 8))
            while (ns > 3 && warp_walk_smem_bytes(ws->xw, 1, ns, slots, want_down, true) > (size_t)ws->smem_optin / 2 - 1024) ns--;
        if (warp_walk_smem_bytes(ws->xw, 1, ns, slots, want_down, true) <= (size_t)ws->smem_optin) {
            s.warp = true; s.wmask = true; s.hybrid = hybrid;
            s.xw = ws->xw; s.xc = 1; s.xns = ns;
            s.n_xtiles = ws->chunk_xtiles[ci].second - ws->chunk_xtiles[ci].first;
            s.a.tiles = ws->d_xtiles.p + ws->chunk_xtiles[ci].first;
            s.a.smem_slots = slots;
            return BNK_OK;
        }
    }
    if (!general && !ws->force_generic && t->max_children <= 2 && K == 20 && ws->walker_mode != 1) {
        const bool hybrid = t->wide_h > 0 && !ws->no_wide;
        const int slots = use_program(hybrid);
        // ring stages: as many as let two CTAs share an SM (the grid is sized to be resident at once)
        int ns = ws->cfg_xns >= 2 && ws->cfg_xns <= 8 ? ws->cfg_xns : 8;
        if (!(ws->cfg_xns >= 2 && ws->cfg_xns <= 8))
            while (ns > 3 && warp_walk_smem_bytes(ws->xw, ws->xc, ns, slots, want_down) > (size_t)ws->smem_optin / 2 - 1024) ns--;
        if (warp_walk_smem_bytes(ws->xw, ws->xc, ns, slots, want_down) <= (size_t)ws->smem_optin) {
            s.fast = true; s.warp = true; s.hybrid = hybrid;
            s.xw = ws->xw; s.xc = ws->xc; s.xns = ns;
            s.n_xtiles = ws->chunk_xtiles[ci].second - ws->chunk_xtiles[ci].first;
            s.a.tiles = ws->d_xtiles.p + ws->chunk_xtiles[ci].first;
            s.a.smem_slots = slots;
            if (ws->walker_mode == 0 || ws->walker_mode == 3) {  // sum-product sweeps: the DMMA variants
                const int mw = std::abs(ws->mw);
                int mns = ws->cfg_xns >= 2 && ws->cfg_xns <= 8 ? ws->cfg_xns : 8;
                if (!(ws->cfg_xns >= 2 && ws->cfg_xns <= 8))
                    while (mns > 3 && mma_walk_smem_bytes(mw, mns, slots, want_down) > (size_t)ws->smem_optin / 2 - 1024) mns--;
                if (mma_walk_smem_bytes(mw, mns, slots, want_down) <= (size_t)ws->smem_optin) {
                    s.mma = true; s.mw = mw; s.mns = mns;
                    s.n_mtiles = ws->chunk_mtiles[ci].second - ws->chunk_mtiles[ci].first;
                    s.mtiles = ws->d_mtiles.p + ws->chunk_mtiles[ci].first;
                }
            }
            return BNK_OK;
        }
    }
    if (!general && !ws->force_generic && t->max_children <= 2 && ck.ncat_max == 1) {
        const bool hybrid = t->wide_h > 0 && !ws->no_wide;
        const int slots = use_program(hybrid);
        const size_t need = fast_smem_bytes(K, ws->cpt, ws->ncg, slots);
        if (need <= (size_t)ws->smem_optin) {
            s.fast = true; s.hybrid = hybrid;
            s.a.smem_slots = slots;
            s.wl.smem_bytes = need;
            return BNK_OK;
        }
    }
    // general kernels (masks, evidence anywhere, any arity); the wide levels still run as level
    // launches when the tree is at most binary (wide_general.cu)
    s.hybrid = t->wide_h > 0 && !ws->no_wide;
    const int slots_needed = use_program(s.hybrid);
    // stack slots: as many as fit in shared memory, the rest spill to global memory
    const size_t base = walk_smem_bytes(K, ws->cpt, ws->ncg, ck.ncat_max, 0);
    const size_t per_slot = sizeof(double) * (size_t)ws->ncg * ws->cpt * K;
    const size_t budget = (size_t)ws->smem_optin;
    if (base > budget) return fail(BNK_ERR_ARG, "tile of %d columns x %d categories does not fit shared memory", ws->tile_cols, ck.ncat_max);
    int smem_slots = (int)std::min<size_t>((size_t)slots_needed, (budget - base) / per_slot);
    s.a.smem_slots = smem_slots;
    if (slots_needed > smem_slots) {
        CUDA_TRY(ws->d_spill.ensure((size_t)(slots_needed - smem_slots) * ws->ncols * K));
        s.a.spill = ws->d_spill.p;
    }
    s.wl.smem_bytes = base + per_slot * smem_slots;
    if (t->max_children >= 32) {  // joint up-sweep: the pairwise-addition queue of big multifurcations lives in global scratch
        const int32_t stride = 2 * t->max_children + 2;
        CUDA_TRY(ws->d_qscratch.ensure((size_t)s.wl.n_tiles * ws->threads * ws->cpt * stride));
        s.a.qscratch = ws->d_qscratch.p;
        s.a.qstride = stride;
    }
    return BNK_OK;
}

// arguments common to the level kernels of one chunk
static WideArgs wide_args(bnk_ws *ws, size_t ci, const uint8_t *obs_s)
{
    WideArgs w;
    std::memset(&w, 0, sizeof(w));
    w.tiles = ws->d_wtiles.p + ws->chunk_wtiles[ci].first;
    w.n_tiles = ws->chunk_wtiles[ci].second - ws->chunk_wtiles[ci].first;
    w.tiles_per_cta = 1;
    {   // site-pattern compression of the height-1 level pays when a group holds enough columns to repeat
        const int g0 = ws->chunk_cgroups[ci].first, ng = ws->chunk_cgroups[ci].second - g0;
        const int cols = ws->chunk_cols[ci].second - ws->chunk_cols[ci].first;
        const int n_h1 = ws->tree->wide_h > 0 ? ws->level_off[1] - ws->level_off[0] : 0;  // every row of the level: the records outlive the launch (cherry-direct)
        if (!ws->no_cherry && ng > 0 && cols / ng >= 64 && n_h1 > 0 &&
            ws->d_cscratch.ensure((size_t)n_h1 * ws->chunk_crecs[ci] * cherry_record_bytes(ws->K)) == cudaSuccess &&
            ws->d_cslots.ensure((size_t)n_h1 * ws->ncols) == cudaSuccess) {
            w.cgroups = ws->d_cgroups.p + g0; w.n_cgroups = ng;
            w.cscratch = ws->d_cscratch.p; w.crecs_per_node = ws->chunk_crecs[ci];
            w.cslots = ws->d_cslots.p;
            w.ctiles = ws->d_ctiles.p + ws->chunk_ctiles[ci].first;
            w.n_ctiles = ws->chunk_ctiles[ci].second - ws->chunk_ctiles[ci].first;
            w.tile_group = ws->d_wtile_group.p + ws->chunk_wtiles[ci].first;
        }
    }
    w.obs = obs_s;
    w.ncols = ws->ncols; w.n_nodes = ws->n;
    w.orig_col = ws->d_orig_col.p;
    w.col_lo = ws->chunk_cols[ci].first; w.col_hi = ws->chunk_cols[ci].second;
    return w;
}

// one launch per wide height level (sliced when a level exceeds the grid's y limit)
template <typename F>
static int for_each_level(bnk_ws *ws, bool ascending, WideArgs &w, F launch, int h_lo = 0, int h_hi = -1)
{
    const int H = ws->tree->wide_h;
    if (h_hi < 0) h_hi = H;
    for (int k = 0; k < H; k++) {
        const int h = ascending ? k : H - 1 - k;
        if (h < h_lo || h >= h_hi) continue;
        for (int off = ws->level_off[h]; off < ws->level_off[h + 1]; off += 65535) {
            const int cnt = std::min(65535, ws->level_off[h + 1] - off);
            w.nodes = ws->d_wide.p + off;
            w.wide_row0 = off;
            w.cherry = (h == 0);
            {   // several tiles per CTA amortise the table loads, as long as ~8 CTAs per SM remain
                const int64_t ctas = (int64_t)cnt * w.n_tiles;
                int tpc = (int)(ctas / ((int64_t)ws->sm_count * 8));
                w.tiles_per_cta = std::max(1, std::min(tpc, ws->wide_tpc_max));
            }
            CUDA_TRY(launch(w, cnt));
            ws->launches++;
        }
    }
    return BNK_OK;
}

// slices of a node list that is not a whole level (the grid's y limit is 65535)
template <typename F>
static int for_node_slices(bnk_ws *ws, WideArgs &w, int count, F launch)
{
    for (int off = 0; off < count; off += 65535) {
        const int cnt = std::min(65535, count - off);
        const int64_t ctas = (int64_t)cnt * w.n_tiles;
        const int tpc = (int)(ctas / ((int64_t)ws->sm_count * 8));
        w.tiles_per_cta = std::max(1, std::min(tpc, ws->wide_tpc_max));
        CUDA_TRY(launch(w, off, cnt));
        ws->launches++;
    }
    return BNK_OK;
}

// Up-sweep over the wide levels, lowest first.  Cherry-direct (lean kernels, site-pattern compression active, the
// fused down-sweep available): the height-1 level only builds its record tables; the nodes under a walked parent
// are expanded into message rows for the walker, everybody else's records are gathered by the parent level
// (wide_common.cuh) and never reach HBM as rows.
static bool cherry_direct(const bnk_ws *ws, const WideArgs &w, bool gen, int K)
{
    return !gen && K == 20 && w.cgroups != nullptr && ws->tree->wide_h >= 2 && ws->n_fuse2 > 0 && !ws->no_fuse2 &&
           !ws->no_cherry_direct;
}
static int wide_up_levels(bnk_ws *ws, WideArgs &w, int K, int mode, bool gen, int n_wt)
{
    auto plain = [&](const WideArgs &wa, int cnt) {
        return gen ? launch_up_wide_gen(K, mode, wa, n_wt, cnt, ws->stream) : launch_up_wide(K, mode, wa, n_wt, cnt, ws->stream); };
    w.cherry_row = nullptr; w.rest_rows = nullptr;
    if (!cherry_direct(ws, w, gen, K)) return for_each_level(ws, true, w, plain);
    TRY(for_each_level(ws, true, w, [&](const WideArgs &wa, int cnt) { return launch_cherry_pairs(K, mode, wa, cnt, ws->stream); }, 0, 1));
    {
        WideArgs r = w;
        TRY(for_node_slices(ws, r, ws->n_wide_rest, [&](WideArgs &wa, int off, int cnt) {
            wa.nodes = ws->d_wide_rest.p + off;
            wa.rest_rows = ws->d_rest_rows.p + off;
            return launch_cherry_expand(K, mode, wa, cnt, ws->stream); }));
    }
    w.cherry_row = ws->d_cherry_row.p;
    return for_each_level(ws, true, w, plain, 1);
}

// --------------------------------------------------------------------------------------
// joint
// --------------------------------------------------------------------------------------
extern "C" int bnk_joint_dev(bnk_ws *ws, int32_t n_cols, const uint8_t *d_obs, const uint8_t *d_present, uint8_t *d_out_state)
{
    if (!ws || !d_obs || !d_out_state) return fail(BNK_ERR_ARG, "joint: null argument");
    ENTER_DEVICE(ws->device);
    TRY(ensure_rates(ws, n_cols));
    ws->twin_phases = false;
    const int K = ws->K, nc = ws->ncols;
    bool general;
    const uint8_t *obs_s, *present_s;
    TRY(prepare_inputs(ws, d_obs, d_present, general, obs_s, present_s));
    if (ws->n_int > 0) {
        CUDA_TRY(ws->d_ptr.ensure((size_t)ws->n_int * nc * K));
        CUDA_TRY(ws->d_kind.ensure((size_t)ws->n_int * nc));
    }
    // states of internal nodes are produced in sorted column space, then scattered (identity: in place)
    uint8_t *state_s = d_out_state;
    if (!ws->identity_perm) { CUDA_TRY(ws->d_state_s.ensure((size_t)ws->n * nc)); state_s = ws->d_state_s.p; }
    for (size_t ci = 0; ci < ws->chunks.size(); ci++) {
        const Chunk &ck = ws->chunks[ci];
        TRY(build_tables(ws, ck, true));
        if (ws->n_int > 0) {
            WalkSetup s;
            TRY(setup_walk(ws, ck, general, false, obs_s, present_s, s));
            s.a.T_row = ws->d_L.p; s.a.T_col = ws->d_LT.p; s.a.prior = ws->md().logprior;
            s.a.ptr = ws->d_ptr.p; s.a.kind = ws->d_kind.p;
            WideArgs w = wide_args(ws, ci, obs_s);
            w.T_row = ws->d_L.p; w.T_col = ws->d_LT.p; w.ptr = ws->d_ptr.p; w.state_s = state_s;
            w.present = present_s; w.prior = ws->md().logprior; w.kind = ws->d_kind.p;
            const int n_wt = ws->chunk_wtiles[ci].second - ws->chunk_wtiles[ci].first;
            const bool gen = !s.fast;
            CUDA_TRY(cudaEventRecord(ws->ev[1][0], ws->stream));
            if (s.hybrid) {
                CUDA_TRY(ws->d_msg.ensure((size_t)ws->n_int * nc * K));
                w.msg = ws->d_msg.p; s.a.msg = ws->d_msg.p;
                TRY(wide_up_levels(ws, w, K, 0, gen, n_wt));
            }
            if (s.wmask) {
                CUDA_TRY(launch_walk_case(s.a.up, s.a.n_steps, present_s, ws->d_orig_col.p, nc, ws->d_kind.p, ws->stream));
                ws->launches++;
                s.a.caseb = ws->d_kind.p;
            }
            CUDA_TRY(s.warp ? launch_joint_up_warp(s.a, s.n_xtiles, s.xw, s.xc, s.xns, ws->stream)
                            : s.fast ? launch_joint_up_fast(s.wl, s.a) : launch_joint_up(s.wl, s.a));
            ws->launches++;
            CUDA_TRY(cudaEventRecord(ws->ev[1][1], ws->stream));
            ws->ev_valid[1] = true;
            // traceback: the walked part first (it holds the roots), then the wide levels top-down
            TracebackArgs ta;
            std::memset(&ta, 0, sizeof(ta));
            ta.down = s.a.down; ta.n_steps = s.a.n_steps; ta.ncols = nc; ta.n_nodes = ws->n; ta.K = K;
            ta.ptr = ws->d_ptr.p; ta.kind = s.fast ? nullptr : ws->d_kind.p;
            ta.kind_obs = ws->masked_only ? 0 : 1;  // evidence on childless nodes only: no walker node is KIND_OBS
            ta.obs = d_obs; ta.orig_col = ws->d_orig_col.p; ta.state_s = state_s; ta.tb_slots = s.tb_slots;
            ta.col_lo = w.col_lo; ta.col_hi = w.col_hi;
            CUDA_TRY(cudaEventRecord(ws->ev[2][0], ws->stream));
            CUDA_TRY(launch_traceback(ta, ws->stream));
            ws->launches++;
            if (s.hybrid) {
                const bool direct = w.cherry_row != nullptr;  // the height-1 level reads its pointers from the record tables
                TRY(for_each_level(ws, false, w, [&](const WideArgs &wa, int cnt) {
                    return gen ? launch_traceback_wide_gen(K, wa, cnt, ws->stream) : launch_traceback_wide(K, wa, cnt, ws->stream); },
                    direct ? 1 : 0));
                if (direct)
                    TRY(for_each_level(ws, false, w, [&](const WideArgs &wa, int cnt) { return launch_traceback_cherry(K, wa, cnt, ws->stream); }, 0, 1));
            }
        } else {
            CUDA_TRY(cudaEventRecord(ws->ev[2][0], ws->stream));
        }
        if (ci + 1 == ws->chunks.size()) {  // all chunks done: back to the caller's column order, then the leaves
            if (!ws->identity_perm && ws->n_int > 0) {
                CUDA_TRY(launch_scatter_rows(state_s, d_out_state, ws->d_node_of_ent.p, ws->n_int, ws->d_orig_col.p, nc, ws->stream));
                ws->launches++;
            }
        }
        CUDA_TRY(cudaEventRecord(ws->ev[2][1], ws->stream));
        ws->ev_valid[2] = true;
    }
    // childless nodes (original column space; may read their parents' states)
    for (size_t ci = 0; ci < ws->chunks.size(); ci++) {
        const Chunk &ck = ws->chunks[ci];
        if (ws->chunks.size() > 1) TRY(build_tables(ws, ck, true));  // the rare uninstantiated-leaf case reads L
        TracebackArgs ta;
        std::memset(&ta, 0, sizeof(ta));
        ta.ncols = nc; ta.n_nodes = ws->n; ta.K = K;
        ta.obs = d_obs; ta.present = d_present; ta.out_state = d_out_state;
        ta.leaf_nodes = ws->d_leaf_nodes.p; ta.leaf_parent = ws->d_leaf_parent.p; ta.n_leaves = ws->n_leaves;
        ta.L = ws->d_L.p;
        ta.cat_of_col = ws->d_cat_of_col.p; ta.cat_lo = ck.cat_lo; ta.cat_hi = ck.cat_hi;
        CUDA_TRY(launch_leaf_states(ta, ws->stream));
        ws->launches++;
    }
    return BNK_OK;
}

extern "C" int bnk_joint(bnk_ws *ws, int32_t n_cols, const uint8_t *obs, const uint8_t *present, uint8_t *out_state)
{
    if (!ws || !obs || !out_state) return fail(BNK_ERR_ARG, "joint: null argument");
    ENTER_DEVICE(ws->device);
    TRY(ensure_rates(ws, n_cols));
    if (ws_wants_pipeline(ws, n_cols)) {
        HostCall hc{0, n_cols, 0, n_cols, obs, present, ws->rates_null ? nullptr : ws->last_rates.data(), out_state, nullptr, 0, nullptr, nullptr, nullptr};
        return ws_run_host_range(ws, hc);
    }
    const size_t bytes = (size_t)ws->n * n_cols;
    CUDA_TRY(ws->d_obs_in.ensure(bytes));
    CUDA_TRY(ws->d_state.ensure(bytes));
    CUDA_TRY(cudaMemcpyAsync(ws->d_obs_in.p, obs, bytes, cudaMemcpyHostToDevice, ws->stream));
    if (present) {
        CUDA_TRY(ws->d_present_in.ensure(bytes));
        CUDA_TRY(cudaMemcpyAsync(ws->d_present_in.p, present, bytes, cudaMemcpyHostToDevice, ws->stream));
    }
    TRY(bnk_joint_dev(ws, n_cols, ws->d_obs_in.p, present ? ws->d_present_in.p : nullptr, ws->d_state.p));
    CUDA_TRY(cudaMemcpyAsync(out_state, ws->d_state.p, bytes, cudaMemcpyDeviceToHost, ws->stream));
    CUDA_TRY(cudaStreamSynchronize(ws->stream));
    return BNK_OK;
}

// --------------------------------------------------------------------------------------
// marginal / log-likelihood
// --------------------------------------------------------------------------------------
int bnk::ws_run_sum(bnk_ws *ws, int32_t n_cols, const uint8_t *d_obs, const uint8_t *d_present, const int32_t *query_nodes,
                   int32_t n_query, double *d_out_post, double *d_out_logl, double *d_out_sum)
{
    ENTER_DEVICE(ws->device);
    TRY(ensure_rates(ws, n_cols));
    const bnk_tree *t = ws->tree;
    const int K = ws->K, nc = ws->ncols;
    const bool want_post = d_out_post != nullptr;
    // query rows.  A node asked for twice is computed once and its row copied; a childless node gets its
    // posterior from its parent's row (leaf_post_kernel), which is computed into an extra row when the
    // parent itself was not asked for.
    std::vector<int32_t> qrow(std::max(ws->n_int, 1), -1);
    struct LeafQuery { int32_t q, node, parent, parent_row; };
    std::vector<LeafQuery> leaf_queries;
    std::vector<std::pair<int32_t, int32_t>> dup_rows;  // (computed row, row to fill)
    int32_t n_rows = n_query < 0 ? ws->n_int : n_query;
    double *post_base = d_out_post;
    if (want_post) {
        if (n_query < 0) {
            for (int32_t e = 0; e < ws->n_int; e++) qrow[e] = e;
        } else {
            if (!query_nodes && n_query > 0) return fail(BNK_ERR_ARG, "marginal: query_nodes is null");
            std::vector<int32_t> leaf_first(t->n, -1);
            for (int32_t q = 0; q < n_query; q++) {
                const int32_t v = query_nodes[q];
                if (v < 0 || v >= t->n) return fail(BNK_ERR_QUERY, "marginal: query node %d is not a node of the tree", v);
                if (t->ent_of_node[v] >= 0) {
                    if (qrow[t->ent_of_node[v]] >= 0) dup_rows.push_back({qrow[t->ent_of_node[v]], q});
                    else qrow[t->ent_of_node[v]] = q;
                } else if (leaf_first[v] >= 0) {
                    dup_rows.push_back({leaf_first[v], q});
                } else {
                    leaf_first[v] = q;
                    leaf_queries.push_back(LeafQuery{q, v, t->parent[v], -1});
                }
            }
            for (LeafQuery &lq : leaf_queries) {
                if (lq.parent < 0) continue;
                int32_t &r = qrow[t->ent_of_node[lq.parent]];
                if (r < 0) r = n_rows++;  // extra row behind the caller's
                lq.parent_row = r;
            }
            if (n_rows > n_query) {  // the sweep writes into a private buffer; the caller's rows are copied out below
                CUDA_TRY(ws->d_post_x.ensure((size_t)n_rows * nc * K));
                post_base = ws->d_post_x.p;
            }
        }
        TRY(upload(ws->d_qrow, qrow, ws->stream));
    }
    bool general;
    const uint8_t *obs_s, *present_s;
    TRY(prepare_inputs(ws, d_obs, d_present, general, obs_s, present_s));
    double *logl = d_out_logl;
    if (!logl && d_out_sum) { CUDA_TRY(ws->d_logl.ensure(nc)); logl = ws->d_logl.p; }
    if (want_post && ws->n_int > 0) {
        CUDA_TRY(ws->d_msg.ensure((size_t)ws->n_int * nc * K));
        if (general) CUDA_TRY(ws->d_kindm.ensure((size_t)ws->n_int * nc));
    }
    if (logl && ws->n_int == 0) { CUDA_TRY(launch_fill_f64(logl, nc, 0.0, ws->stream)); ws->launches++; }
    for (size_t ci = 0; ci < ws->chunks.size() && ws->n_int > 0; ci++) {
        const Chunk &ck = ws->chunks[ci];
        TRY(build_tables(ws, ck, false));
        WalkSetup s;
        TRY(setup_walk(ws, ck, general, want_post, obs_s, present_s, s));
        s.a.T_row = ws->d_P.p; s.a.T_col = ws->d_PT.p; s.a.prior = ws->md().prior;
        s.a.msg = want_post ? ws->d_msg.p : nullptr;
        s.a.kindm = (want_post && general) ? ws->d_kindm.p : nullptr;
        s.a.out_logl = logl;
        WideArgs w = wide_args(ws, ci, obs_s);
        w.T_row = ws->d_P.p; w.T_col = ws->d_PT.p;
        w.present = present_s; w.prior = ws->md().prior;
        const int n_wt = ws->chunk_wtiles[ci].second - ws->chunk_wtiles[ci].first;
        const bool gen = !s.fast;
        CUDA_TRY(cudaEventRecord(ws->ev[3][0], ws->stream));
        if (s.hybrid) {
            CUDA_TRY(ws->d_msg.ensure((size_t)ws->n_int * nc * K));  // the walker reads the wide messages even for logL only
            w.msg = ws->d_msg.p; s.a.msg = ws->d_msg.p;
            if (gen) {
                CUDA_TRY(ws->d_kindm.ensure((size_t)ws->n_int * nc));
                w.kindm = ws->d_kindm.p; s.a.kindm = ws->d_kindm.p;
            }
            if (logl) {
                CUDA_TRY(ws->d_escale.ensure((size_t)std::max(t->n_wide, 1) * nc));
                CUDA_TRY(ws->d_esum_wide.ensure(nc));
                w.escale = ws->d_escale.p;
                if (gen) {  // component roots and scalar factors inside the wide part
                    CUDA_TRY(ws->d_logl_wide.ensure(nc));
                    if (ci == 0) CUDA_TRY(cudaMemsetAsync(ws->d_logl_wide.p, 0, sizeof(double) * nc, ws->stream));
                    w.logl_acc = ws->d_logl_wide.p;
                    s.a.logl_wide = ws->d_logl_wide.p;
                }
            }
            TRY(wide_up_levels(ws, w, K, 1, gen, n_wt));
            if (logl) {
                CUDA_TRY(launch_sum_escale(ws->d_escale.p, t->n_wide, nc, ws->d_esum_wide.p, ws->stream));
                ws->launches++;
                s.a.esum_wide = ws->d_esum_wide.p;
            }
        }
        if (s.mma) s.a.tiles = s.mtiles;
        if (s.wmask) {
            CUDA_TRY(ws->d_kindm.ensure((size_t)ws->n_int * nc));
            CUDA_TRY(launch_walk_case(s.a.up, s.a.n_steps, present_s, ws->d_orig_col.p, nc, ws->d_kindm.p, ws->stream));
            ws->launches++;
            s.a.caseb = ws->d_kindm.p;
        }
        CUDA_TRY(s.mma ? launch_sum_up_mma(s.a, s.n_mtiles, s.mw, s.mns, ws->stream)
                 : s.warp ? launch_sum_up_warp(s.a, s.n_xtiles, s.xw, s.xc, s.xns, ws->stream)
                          : s.fast ? launch_sum_up_fast(s.wl, s.a) : launch_sum_up(s.wl, s.a));
        ws->launches++;
        CUDA_TRY(cudaEventRecord(ws->ev[3][1], ws->stream));
        ws->ev_valid[3] = true;
        if (want_post) {
            s.a.T_row = ws->d_PT.p;
            s.a.T_aux = ws->d_P.p;
            s.a.out_post = post_base;
            s.a.qrow = (n_query < 0) ? nullptr : ws->d_qrow.p;
            if (s.hybrid) {
                CUDA_TRY(ws->d_tmsg.ensure((size_t)ws->n_int * nc * K));
                s.a.tmsg = ws->d_tmsg.p;
            }
            CUDA_TRY(cudaEventRecord(ws->ev[4][0], ws->stream));
            if (s.mma) s.a.tiles = ws->d_xtiles.p + ws->chunk_xtiles[ci].first;
            CUDA_TRY(s.warp ? launch_sum_down_warp(s.a, s.n_xtiles, s.xw, s.xc, s.xns, ws->stream)
                              : s.fast ? launch_sum_down_fast(s.wl, s.a) : launch_sum_down(s.wl, s.a));
            ws->launches++;
            if (s.hybrid) {
                w.tmsg = ws->d_tmsg.p; w.out_post = post_base; w.qrow = s.a.qrow;
                const bool fuse2 = !gen && K == 20 && t->wide_h >= 2 && ws->n_fuse2 > 0 && !ws->no_fuse2;
                if (!fuse2)
                    TRY(for_each_level(ws, false, w, [&](const WideArgs &wa, int cnt) {
                        return gen ? launch_down_wide_gen(K, wa, n_wt, cnt, ws->stream) : launch_down_wide(K, wa, n_wt, cnt, ws->stream); }));
                if (fuse2) {
                    // every wide node above height 1 evaluates its height-1 children in the same thread (their
                    // from-above vectors never reach HBM); height-1 nodes under a walked parent keep the plain kernel
                    w.cherry = 0;
                    for (int h = t->wide_h - 1; h >= 1; h--) {
                        const int base = ws->fuse_off[h];
                        TRY(for_node_slices(ws, w, ws->fuse_off[h + 1] - base, [&](WideArgs &wa, int off, int cnt) {
                            wa.fuse2 = ws->d_fuse2.p + base + off;
                            return launch_down_wide2(K, wa, n_wt, cnt, ws->stream); }));
                    }
                    w.fuse2 = nullptr;
                    w.cherry = 1;
                    TRY(for_node_slices(ws, w, ws->n_wide_rest, [&](WideArgs &wa, int off, int cnt) {
                        wa.nodes = ws->d_wide_rest.p + off;
                        return launch_down_wide(K, wa, n_wt, cnt, ws->stream); }));
                }
            }
            for (const LeafQuery &lq : leaf_queries) {
                LeafPostArgs la;
                std::memset(&la, 0, sizeof(la));
                la.node = lq.node; la.parent = lq.parent; la.ncols = nc; la.n_nodes = ws->n; la.K = K;
                la.obs = d_obs; la.present = d_present; la.P = ws->d_P.p;
                la.cat_of_col = ws->d_cat_of_col.p; la.cat_lo = ck.cat_lo; la.cat_hi = ck.cat_hi;
                la.parent_row = lq.parent_row >= 0 ? post_base + (size_t)lq.parent_row * nc * K : nullptr;
                la.out_row = post_base + (size_t)lq.q * nc * K;
                CUDA_TRY(launch_leaf_post(la, ws->stream));
                ws->launches++;
            }
            CUDA_TRY(cudaEventRecord(ws->ev[4][1], ws->stream));
            ws->ev_valid[4] = true;
        }
    }
    if (want_post && ws->n_int == 0)  // a forest of single nodes: nothing has a BN variable
        for (const LeafQuery &lq : leaf_queries) {
            CUDA_TRY(launch_fill_f64(post_base + (size_t)lq.q * nc * K, (size_t)nc * K, NAN, ws->stream));
            ws->launches++;
        }
    if (want_post && post_base != d_out_post && n_query > 0)
        CUDA_TRY(cudaMemcpyAsync(d_out_post, post_base, sizeof(double) * (size_t)n_query * nc * K, cudaMemcpyDeviceToDevice, ws->stream));
    for (const auto &dr : dup_rows)
        CUDA_TRY(cudaMemcpyAsync(d_out_post + (size_t)dr.second * nc * K, d_out_post + (size_t)dr.first * nc * K,
                                 sizeof(double) * (size_t)nc * K, cudaMemcpyDeviceToDevice, ws->stream));
    if (d_out_sum) { CUDA_TRY(launch_sum_f64(logl, nc, d_out_sum, ws->stream)); ws->launches++; }
    return BNK_OK;
}

extern "C" int bnk_marginal_dev(bnk_ws *ws, int32_t n_cols, const uint8_t *d_obs, const uint8_t *d_present,
                                const int32_t *query_nodes_host, int32_t n_query, double *d_out_post, double *d_out_logl)
{
    if (!ws || !d_obs) return fail(BNK_ERR_ARG, "marginal: null argument");
    return ws_run_sum(ws, n_cols, d_obs, d_present, query_nodes_host, n_query, d_out_post, d_out_logl, nullptr);
}

// Joint + marginal of the same alignment as ONE call: the two passes share nothing but their inputs (log-space
// tables, pointers and states on one side; linear tables, messages and posteriors on the other), so the joint
// pass runs in a twin workspace on a second stream while this workspace runs the marginal pass.  On deep trees
// both walks are latency-bound with about one warp per scheduler (profiles/r1_v22_ncu_summary.md), and two
// co-resident walker CTAs per SM fill each other's stalls; on level-scheduled trees the issue-bound joint
// kernels and the HBM-bound down-sweep overlap at the tails of their launches.
extern "C" int bnk_joint_marginal_dev(bnk_ws *ws, int32_t n_cols, const uint8_t *d_obs, const uint8_t *d_present,
                                      uint8_t *d_out_state, const int32_t *query_nodes_host, int32_t n_query,
                                      double *d_out_post, double *d_out_logl)
{
    if (!ws || !d_obs || !d_out_state) return fail(BNK_ERR_ARG, "joint_marginal: null argument");
    if (ws->is_sub) return fail(BNK_ERR_STATE, "joint_marginal: not on a sub-workspace");
    ENTER_DEVICE(ws->device);
    TRY(ensure_rates(ws, n_cols));
    {   // Two streams pay where the walks dominate (deep trees: latency-bound kernels fill each other's stalls).  On a
        // level-scheduled tree the kernels of one pass fill the machine on their own, the second stream only adds a
        // second table build and tail effects (C3a: 16.8 ms against 16.4 ms one pass after the other): run in order.
        const bnk_tree *t = ws->tree;
        const bool walker_tree = t->wide_h == 0 || ws->no_wide || 2 * (int64_t)t->narrow.up.size() > t->n_int;
        if (!walker_tree) {
            TRY(bnk_joint_dev(ws, n_cols, d_obs, d_present, d_out_state));
            return ws_run_sum(ws, n_cols, d_obs, d_present, query_nodes_host, n_query, d_out_post, d_out_logl, nullptr);
        }
    }
    if (!ws->twin) {
        bnk_ws *tw = nullptr;
        TRY(bnk_ws_create(&ws->model, ws->tree, ws->max_cols, ws->device, nullptr, &tw));
        tw->is_sub = true;
        ws->twin = tw;
        ws->twin_dist_version = 0;
        CUDA_TRY(cudaEventCreateWithFlags(&ws->twin_fork, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ws->twin_join, cudaEventDisableTiming));
    }
    bnk_ws *tw = ws->twin;
    if (tw->cfg_tile_cols != ws->cfg_tile_cols || tw->cfg_cpt != ws->cfg_cpt || tw->force_generic != ws->force_generic ||
        tw->no_wide != ws->no_wide) {
        tw->cfg_tile_cols = ws->cfg_tile_cols; tw->cfg_cpt = ws->cfg_cpt; tw->force_generic = ws->force_generic;
        tw->no_wide = ws->no_wide; tw->retile = true;
    }
    tw->evidence_mode = ws->evidence_mode;
    if (ws->twin_dist_version != ws->dist_version + 1) TRY(bnk_set_distances(tw, ws->dist.data()));
    ws->twin_dist_version = ws->dist_version + 1;
    TRY(bnk_set_rates(tw, ws->ncols, ws->rates_null ? nullptr : ws->last_rates.data()));
    // fork: the twin's stream sees everything queued on this workspace's stream so far (the inputs)
    CUDA_TRY(cudaEventRecord(ws->twin_fork, ws->stream));
    CUDA_TRY(cudaStreamWaitEvent(tw->stream, ws->twin_fork, 0));
    int rc = bnk_joint_dev(tw, n_cols, d_obs, d_present, d_out_state);
    if (rc == BNK_OK) rc = ws_run_sum(ws, n_cols, d_obs, d_present, query_nodes_host, n_query, d_out_post, d_out_logl, nullptr);
    // join (also after an error: whatever the twin queued must not outlive the call's ordering)
    CUDA_TRY(cudaEventRecord(ws->twin_join, tw->stream));
    CUDA_TRY(cudaStreamWaitEvent(ws->stream, ws->twin_join, 0));
    ws->twin_phases = rc == BNK_OK;
    return rc;
}

extern "C" int bnk_loglik_dev(bnk_ws *ws, int32_t n_cols, const uint8_t *d_obs, const uint8_t *d_present,
                              double *d_out_logl, double *d_out_sum)
{
    if (!ws || !d_obs) return fail(BNK_ERR_ARG, "loglik: null argument");
    return ws_run_sum(ws, n_cols, d_obs, d_present, nullptr, 0, nullptr, d_out_logl, d_out_sum);
}

static int stage_inputs(bnk_ws *ws, int32_t n_cols, const uint8_t *obs, const uint8_t *present)
{
    const size_t bytes = (size_t)ws->n * n_cols;
    CUDA_TRY(ws->d_obs_in.ensure(bytes));
    CUDA_TRY(cudaMemcpyAsync(ws->d_obs_in.p, obs, bytes, cudaMemcpyHostToDevice, ws->stream));
    if (present) {
        CUDA_TRY(ws->d_present_in.ensure(bytes));
        CUDA_TRY(cudaMemcpyAsync(ws->d_present_in.p, present, bytes, cudaMemcpyHostToDevice, ws->stream));
    }
    return BNK_OK;
}

extern "C" int bnk_marginal(bnk_ws *ws, int32_t n_cols, const uint8_t *obs, const uint8_t *present,
                            const int32_t *query_nodes, int32_t n_query, double *out_post, double *out_logl)
{
    if (!ws || !obs) return fail(BNK_ERR_ARG, "marginal: null argument");
    ENTER_DEVICE(ws->device);
    TRY(ensure_rates(ws, n_cols));
    if (ws_wants_pipeline(ws, n_cols)) {
        HostCall hc{1, n_cols, 0, n_cols, obs, present, ws->rates_null ? nullptr : ws->last_rates.data(), nullptr, query_nodes, n_query, out_post, out_logl, nullptr};
        return ws_run_host_range(ws, hc);
    }
    TRY(stage_inputs(ws, n_cols, obs, present));
    const int nq = n_query < 0 ? ws->n_int : n_query;
    const size_t post_elems = (size_t)nq * n_cols * ws->K;
    if (out_post && post_elems) CUDA_TRY(ws->d_post.ensure(post_elems));
    if (out_logl) CUDA_TRY(ws->d_logl.ensure(n_cols));
    TRY(ws_run_sum(ws, n_cols, ws->d_obs_in.p, present ? ws->d_present_in.p : nullptr, query_nodes, n_query,
                (out_post && post_elems) ? ws->d_post.p : nullptr, out_logl ? ws->d_logl.p : nullptr, nullptr));
    if (out_post && post_elems)
        CUDA_TRY(cudaMemcpyAsync(out_post, ws->d_post.p, post_elems * sizeof(double), cudaMemcpyDeviceToHost, ws->stream));
    if (out_logl) CUDA_TRY(cudaMemcpyAsync(out_logl, ws->d_logl.p, sizeof(double) * n_cols, cudaMemcpyDeviceToHost, ws->stream));
    CUDA_TRY(cudaStreamSynchronize(ws->stream));
    return BNK_OK;
}

// Joint + marginal with host buffers, one call: column chunks through the two sub-workspaces as in bnk_marginal, each
// chunk running its joint pass too -- the whole joint reconstruction hides under the marginal's result copies (PCIe).
extern "C" int bnk_joint_marginal(bnk_ws *ws, int32_t n_cols, const uint8_t *obs, const uint8_t *present, uint8_t *out_state,
                                  const int32_t *query_nodes, int32_t n_query, double *out_post, double *out_logl)
{
    if (!ws || !obs || !out_state) return fail(BNK_ERR_ARG, "joint_marginal: null argument");
    ENTER_DEVICE(ws->device);
    TRY(ensure_rates(ws, n_cols));
    if (ws_wants_pipeline(ws, n_cols)) {
        HostCall hc{3, n_cols, 0, n_cols, obs, present, ws->rates_null ? nullptr : ws->last_rates.data(), out_state, query_nodes, n_query, out_post, out_logl, nullptr};
        return ws_run_host_range(ws, hc);
    }
    TRY(bnk_joint(ws, n_cols, obs, present, out_state));
    return bnk_marginal(ws, n_cols, obs, present, query_nodes, n_query, out_post, out_logl);
}

extern "C" int bnk_loglik(bnk_ws *ws, int32_t n_cols, const uint8_t *obs, const uint8_t *present, double *out_logl, double *out_sum)
{
    if (!ws || !obs) return fail(BNK_ERR_ARG, "loglik: null argument");
    ENTER_DEVICE(ws->device);
    TRY(ensure_rates(ws, n_cols));
    if (ws_wants_pipeline(ws, n_cols)) {
        double total = 0.0;
        HostCall hc{2, n_cols, 0, n_cols, obs, present, ws->rates_null ? nullptr : ws->last_rates.data(), nullptr, nullptr, 0, nullptr, out_logl, &total};
        TRY(ws_run_host_range(ws, hc));
        if (out_sum) *out_sum = total;
        return BNK_OK;
    }
    TRY(stage_inputs(ws, n_cols, obs, present));
    CUDA_TRY(ws->d_logl.ensure(n_cols));
    TRY(ws_run_sum(ws, n_cols, ws->d_obs_in.p, present ? ws->d_present_in.p : nullptr, nullptr, 0, nullptr, ws->d_logl.p,
                out_sum ? ws->d_sum.p : nullptr));
    if (out_logl) CUDA_TRY(cudaMemcpyAsync(out_logl, ws->d_logl.p, sizeof(double) * n_cols, cudaMemcpyDeviceToHost, ws->stream));
    if (out_sum) CUDA_TRY(cudaMemcpyAsync(out_sum, ws->d_sum.p, sizeof(double), cudaMemcpyDeviceToHost, ws->stream));
    CUDA_TRY(cudaStreamSynchronize(ws->stream));
    return BNK_OK;
}
