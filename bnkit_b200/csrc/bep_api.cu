// bep_api.cu -- bnk_bep_presence: the per-column presence mask of every ancestor by bi-directional edge
// parsimony, as asr.Prediction.PredictByBidirEdgeParsimony + patchAncestorsWithBEP produce it
// (src/asr/Prediction.java:1443-1590) and Prediction.getTree consumes it (:336-349).
//
// Device: the parsimony sweeps of all 2 (nPos + 1) edge instances (bep.cu).  Host: what the reference does per
// ancestor after inference -- assemble the partial-order graph from the optimal edges (POGraph.createFromEdgeMap,
// src/dat/pog/POGraph.java:544-559), test it for a start-to-end path (isContiguous :314-335), remove nodes that
// stick out (nibble :566-617), and for graphs without a path re-infer the protruding columns without the NULL
// symbol until a path exists (patchAncestorsWithBEP; those re-inferences run on the device again).
#include "bep.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <set>
#include <thread>
#include <vector>

using namespace bnk;

#define BEP_CUDA_TRY(expr)                                                                                  \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess) {                                                                            \
            rc = fail(e_ == cudaErrorMemoryAllocation ? BNK_ERR_NOMEM : BNK_ERR_CUDA, "%s: %s (%s:%d)", #expr, \
                      cudaGetErrorString(e_), __FILE__, __LINE__);                                          \
            goto done;                                                                                      \
        }                                                                                                   \
    } while (0)

// the ancestors' partial-order graphs as bnk_bep_infer hands them out (<-> Map<Object, POGraph> of asr.Prediction)
struct bnk_pogs {
    int32_t n_pos = 0;
    std::vector<int32_t> ent_of_node;
    std::vector<std::vector<int32_t>> from, to;   // per ancestor row: edges sorted by (from, to); -1 = start, n_pos = end
    std::vector<std::vector<uint8_t>> flags;      // bit 0: inferred in the forward direction, bit 1: backward (POGraph.BidirEdge)
};

namespace {

// an inferred edge: (from, to, direction flags)
struct Edge3 {
    int a, b;
    uint8_t f;
    bool operator<(const Edge3 &o) const { return a != o.a ? a < o.a : b < o.b; }
};
// sort, and merge duplicates by OR-ing their direction flags (EdgeMap.Directed.add, src/dat/pog/EdgeMap.java)
inline void merge_edges(std::vector<Edge3> &e)
{
    std::sort(e.begin(), e.end());
    size_t w = 0;
    for (size_t r = 0; r < e.size(); r++) {
        if (w > 0 && e[w - 1].a == e[r].a && e[w - 1].b == e[r].b) e[w - 1].f |= e[r].f;
        else e[w++] = e[r];
    }
    e.resize(w);
}

// dat.pog.POGraph restricted to what BEP needs: directed, terminated, nodes = alignment columns
struct Pog {
    int n = 0;
    std::vector<char> node;
    std::vector<std::vector<int>> fwd, bwd;
    std::vector<char> is_start, is_end;
    std::vector<int> touched;  // nodes ever added (for cheap reuse)

    void reset(int n_pos)
    {
        if (n != n_pos) {
            n = n_pos;
            node.assign(n, 0); fwd.assign(n, {}); bwd.assign(n, {}); is_start.assign(n, 0); is_end.assign(n, 0);
            touched.clear();
            return;
        }
        for (int i : touched) { node[i] = 0; fwd[i].clear(); bwd[i].clear(); is_start[i] = 0; is_end[i] = 0; }
        touched.clear();
    }
    bool is_node(int i) const { return i >= 0 && i < n && node[i]; }
    void add_node(int i)  // IdxGraph.addNode (src/dat/pog/IdxGraph.java:346-357)
    {
        node[i] = 1; fwd[i].clear(); bwd[i].clear();
        touched.push_back(i);
    }
    static bool has(const std::vector<int> &v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); }
    bool add_edge(int a, int b)  // IdxGraph.addEdge (:402-419); returns whether the graph changed
    {
        if (is_node(a) && is_node(b)) {
            if (has(fwd[a], b)) return false;
            fwd[a].push_back(b); bwd[b].push_back(a);
            return true;
        }
        if (a == -1 && is_node(b)) { const bool c = !is_start[b]; is_start[b] = 1; return c; }
        if (is_node(a) && b == n) { const bool c = !is_end[a]; is_end[a] = 1; return c; }
        return false;
    }
    void remove_node(int i)  // IdxGraph.removeNode (:373-391)
    {
        for (int nx : fwd[i]) { auto &v = bwd[nx]; v.erase(std::remove(v.begin(), v.end(), i), v.end()); }
        for (int pv : bwd[i]) { auto &v = fwd[pv]; v.erase(std::remove(v.begin(), v.end(), i), v.end()); }
        node[i] = 0; fwd[i].clear(); bwd[i].clear(); is_start[i] = 0; is_end[i] = 0;
    }
    std::vector<int> order() const  // ascending = a topological order (every edge runs to a higher column)
    {
        std::vector<int> o;
        for (int i : touched) if (node[i]) o.push_back(i);
        std::sort(o.begin(), o.end());
        o.erase(std::unique(o.begin(), o.end()), o.end());
        return o;
    }
    bool is_contiguous() const  // isPath(-1, N)
    {
        std::vector<int> stack;
        std::vector<char> seen(n, 0);
        for (int i : touched) if (node[i] && is_start[i] && !seen[i]) { seen[i] = 1; stack.push_back(i); }
        while (!stack.empty()) {
            const int v = stack.back(); stack.pop_back();
            if (is_end[v]) return true;
            for (int nx : fwd[v]) if (!seen[nx]) { seen[nx] = 1; stack.push_back(nx); }
        }
        return false;
    }
    bool sticks_out(int i, bool from_start) const  // POGraph.nibble(index, direction) (:566-581)
    {
        if (from_start) return !is_start[i] && bwd[i].empty();
        return !is_end[i] && fwd[i].empty();
    }
    void nibble()  // (:588-617)
    {
        const std::vector<int> o = order();
        for (int i : o) if (node[i] && sticks_out(i, true)) remove_node(i);
        for (auto it = o.rbegin(); it != o.rend(); ++it) if (node[*it] && sticks_out(*it, false)) remove_node(*it);
    }
    std::set<int> protruding() const  // getProtruding (:622-646): +i sticks out towards the end, -i towards the start
    {
        std::set<int> cols;
        const std::vector<int> o = order();
        for (int i : o) if (sticks_out(i, false)) cols.insert(i);
        for (auto it = o.rbegin(); it != o.rend(); ++it) if (sticks_out(*it, true)) cols.insert(-*it);
        return cols;
    }
    void from_edges(int n_pos, const std::vector<Edge3> &edges)  // createFromEdgeMap (:544-559)
    {
        reset(n_pos);
        for (const auto &ed : edges) {
            if (ed.a != -1 && !is_node(ed.a)) add_node(ed.a);
            if (ed.b != n_pos && !is_node(ed.b)) add_node(ed.b);
            add_edge(ed.a, ed.b);
        }
    }
    // the graph as it stands, with the direction flags of `edges` (every edge of the graph is in that list)
    void export_to(bnk_pogs *out, int ent, const std::vector<Edge3> &edges) const
    {
        auto flag_of = [&](int a, int b) -> uint8_t {
            for (const Edge3 &e : edges) if (e.a == a && e.b == b) return e.f;   // lists are short; patched lists are unsorted
            return 0;
        };
        std::vector<int32_t> &fr = out->from[ent], &to = out->to[ent];
        std::vector<uint8_t> &fl = out->flags[ent];
        fr.clear(); to.clear(); fl.clear();
        const std::vector<int> o = order();
        for (int i : o) if (is_start[i]) { fr.push_back(-1); to.push_back(i); fl.push_back(flag_of(-1, i)); }
        for (int i : o) {
            std::vector<int> nx = fwd[i];
            std::sort(nx.begin(), nx.end());
            for (int j : nx) { fr.push_back(i); to.push_back(j); fl.push_back(flag_of(i, j)); }
            if (is_end[i]) { fr.push_back(i); to.push_back(n); fl.push_back(flag_of(i, n)); }
        }
    }
};

// Ancestors are independent after inference (the reference's own TODO: "use threads here too",
// Prediction.java:1483): static slices of the work list over the host's hardware threads.
template <typename F>
void parallel_slices(int n_items, F fn)
{
    int nt = (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 32) nt = 32;
    if (n_items < 64) nt = 1;
    if (nt == 1) { fn(0, n_items); return; }
    std::vector<std::thread> th;
    for (int k = 0; k < nt; k++) {
        const int lo = (int)((int64_t)n_items * k / nt), hi = (int)((int64_t)n_items * (k + 1) / nt);
        if (hi > lo) th.emplace_back([=] { fn(lo, hi); });
    }
    for (auto &x : th) x.join();
}

// a device buffer that keeps its allocation between calls (alloc = "make room for n elements")
template <typename T>
struct Dev {
    T *p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t alloc(size_t n)
    {
        n = std::max<size_t>(n, 1);
        if (n <= cap) return cudaSuccess;
        drop();
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = n;
        return cudaSuccess;
    }
    void drop() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    void trim(size_t max_bytes) { if (cap * sizeof(T) > max_bytes) drop(); }
};

// Per-device scratch that survives the call: stream, events and every buffer up to KEEP_BYTES.  At the sizes of the
// reference's own data (C1: 66 x 5, C2: 39 x 1,307) a call used to be all fixed cost -- ~20 cudaMalloc, a stream,
// page-locked staging -- and that cost is now paid once per process; large alignments give their memory back.
struct Pinned {  // page-locked staging for device->host copies, kept between calls up to KEEP_BYTES
    void *p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void trim(size_t max_bytes) { if (cap > max_bytes) { cudaFreeHost(p); p = nullptr; cap = 0; } }
};

struct BepScratch {
    std::mutex mu;
    Pinned pin_dict, pin_opt;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    Dev<uint8_t> d_obs, d_fwd, d_symmap, d_contig, d_present;
    Dev<int32_t> d_leaf_nodes, d_leaf_of_node, d_ent, d_parent, d_first, d_kids, d_nxt, d_prv, d_inst_e, d_dict, d_nsym, d_flag, d_level_nodes, d_rows;
    Dev<uint32_t> d_A, d_B, d_A1, d_Z, d_compact;
    static constexpr size_t KEEP_BYTES = (size_t)64 << 20;
    void trim()
    {
        d_obs.trim(KEEP_BYTES); d_fwd.trim(KEEP_BYTES); d_symmap.trim(KEEP_BYTES); d_contig.trim(KEEP_BYTES); d_present.trim(KEEP_BYTES);
        d_leaf_nodes.trim(KEEP_BYTES); d_leaf_of_node.trim(KEEP_BYTES); d_ent.trim(KEEP_BYTES); d_parent.trim(KEEP_BYTES);
        d_first.trim(KEEP_BYTES); d_kids.trim(KEEP_BYTES); d_nxt.trim(KEEP_BYTES); d_prv.trim(KEEP_BYTES); d_inst_e.trim(KEEP_BYTES);
        d_dict.trim(KEEP_BYTES); d_nsym.trim(KEEP_BYTES); d_flag.trim(KEEP_BYTES); d_level_nodes.trim(KEEP_BYTES); d_rows.trim(KEEP_BYTES);
        d_A.trim(KEEP_BYTES); d_B.trim(KEEP_BYTES); d_A1.trim(KEEP_BYTES); d_Z.trim(KEEP_BYTES); d_compact.trim(KEEP_BYTES);
        pin_dict.trim(KEEP_BYTES); pin_opt.trim(KEEP_BYTES);
    }
};
BepScratch &scratch_of(int device)
{
    static BepScratch pool[64];   // never destroyed: the CUDA context may already be gone at exit
    return pool[device & 63];
}

}  // namespace

static int bep_run(const bnk_tree *t, int32_t n_cols, const uint8_t *obs, int device, uint8_t *present_out,
                   int32_t *info_or_null, bnk_pogs *pogs)
{
    if (!t || !obs || !present_out || n_cols < 1) return fail(BNK_ERR_ARG, "bep: bad argument");
    if (t->root_count != 1 || t->parent[0] >= 0)
        return fail(BNK_ERR_TREE, "bep: the tree must have a single root at index 0 (asr.Parsimony starts at branch point 0)");
    if (t->max_children >= (1 << 20)) return fail(BNK_ERR_TREE, "bep: at most 2^20 - 1 children per node (the tree has a node with %d)", t->max_children);
    int rc = BNK_OK;
    const int n = t->n, n_pos = n_cols, n_int = t->n_int;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(BNK_ERR_CUDA, "bep: no CUDA device visible (libbnkgpu has no CPU fallback)");
    }
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return fail(BNK_ERR_CUDA, "bep: cannot select device %d", device);
    int cur_dev = 0;
    if (cudaGetDevice(&cur_dev) != cudaSuccess) cur_dev = 0;
    BepScratch &S = scratch_of(cur_dev);
    std::lock_guard<std::mutex> scratch_lock(S.mu);   // one BEP call per device at a time
    std::vector<int32_t> leaf_nodes, leaf_of_node(n, -1);
    for (int v = 0; v < n; v++)
        if (t->ent_of_node[v] < 0) { leaf_of_node[v] = (int32_t)leaf_nodes.size(); leaf_nodes.push_back(v); }
    const int n_leaves = (int)leaf_nodes.size();
    // internal nodes grouped by height (1 = all children childless): the level schedule of the sweeps
    int n_levels = 0;
    for (int v = 0; v < n; v++) if (t->ent_of_node[v] >= 0) n_levels = std::max(n_levels, (int)t->height[v]);
    std::vector<int32_t> level_off(n_levels + 1, 0), level_nodes(t->n_int);
    for (int v = 0; v < n; v++) if (t->ent_of_node[v] >= 0) level_off[t->height[v]]++;
    for (int h = 0; h < n_levels; h++) level_off[h + 1] += level_off[h];
    {
        std::vector<int32_t> fill(level_off.begin(), level_off.end() - 1);
        for (int v = 0; v < n; v++) if (t->ent_of_node[v] >= 0) level_nodes[fill[t->height[v] - 1]++] = v;
    }
    // one launch per level pays while levels are wide; a very deep tree (caterpillar) keeps one thread per instance
    bool by_levels = n_levels <= 4096;
    if (const char *env = std::getenv("BNK_BEP_SCHEDULE")) by_levels = std::strcmp(env, "instance") != 0;
    // Graph assembly: on the device (bep_reach.cu) unless the caller wants the graphs themselves (bnk_bep_infer) or
    // BNK_BEP_ASSEMBLY=host asks for the r1 path (A/B runs, tests)
    bool device_assembly = pogs == nullptr;
    if (const char *env = std::getenv("BNK_BEP_ASSEMBLY")) device_assembly = device_assembly && std::strcmp(env, "host") != 0;
    if (!device_assembly || n_int == 0) {
        // extant rows pass through; ancestors are filled in below
        for (int v = 0; v < n; v++) {
            uint8_t *row = present_out + (size_t)v * n_pos;
            if (t->ent_of_node[v] < 0) for (int c = 0; c < n_pos; c++) row[c] = obs[(size_t)v * n_pos + c] != BNK_NONE;
            else std::memset(row, 0, n_pos);
        }
    }
    if (n_int == 0) return BNK_OK;

    cudaStream_t &st = S.st;
    cudaEvent_t &ev0 = S.ev0, &ev1 = S.ev1;
    Dev<uint8_t> &d_obs = S.d_obs, &d_fwd = S.d_fwd, &d_symmap = S.d_symmap, &d_contig = S.d_contig, &d_present = S.d_present;
    Dev<int32_t> &d_leaf_nodes = S.d_leaf_nodes, &d_leaf_of_node = S.d_leaf_of_node, &d_ent = S.d_ent, &d_parent = S.d_parent,
                 &d_first = S.d_first, &d_kids = S.d_kids, &d_nxt = S.d_nxt, &d_prv = S.d_prv, &d_inst_e = S.d_inst_e, &d_dict = S.d_dict,
                 &d_nsym = S.d_nsym, &d_flag = S.d_flag, &d_level_nodes = S.d_level_nodes, &d_rows = S.d_rows;
    Dev<uint32_t> &d_A = S.d_A, &d_B = S.d_B, &d_A1 = S.d_A1, &d_Z = S.d_Z, &d_compact = S.d_compact;
    std::vector<int32_t> opt_row_of;      // device assembly: ancestor row -> row of the compact copy of its optimal sets
    const int MAXSYM = 255;
    // crippled ancestors after the first pass, patch rounds, largest symbol count, kernel launches,
    // microseconds: device sweeps (incl. their copies), host graph assembly, whole call, kernels alone (CUDA events)
    int info[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const auto t_begin = std::chrono::steady_clock::now();
    auto us_since = [](std::chrono::steady_clock::time_point t0) {
        return (int)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
    };
    const bool trace = std::getenv("BNK_BEP_TRACE") != nullptr;  // stderr: microseconds since entry at a few marks
    auto mark = [&](const char *what) { if (trace) std::fprintf(stderr, "bep %8d us  %s\n", us_since(t_begin), what); };
    Pog pog;  // scratch for the serial tail
    // ancestors without a start-to-end path: their edges so far (the graph is rebuilt in `pog` when needed) and
    // the columns that stick out
    std::map<int, std::vector<Edge3>> crippled_edges;
    std::map<int, std::set<int>> crippled;
    std::vector<int32_t> h_nsym;
    Pinned &pin_dict = S.pin_dict, &pin_opt = S.pin_opt;
    const int32_t *h_dict = nullptr;
    const uint32_t *h_opt = nullptr;  // optimal sets of ancestor rows opt_ent0 .. (a chunk, or all rows)
    int opt_ent0 = 0;

    // one batch of instances through the device; results land in h_nsym / h_dict / h_opt ([ent][W][n_inst])
    std::vector<int32_t> want_rows;       // patch rounds: the ancestor rows (ents) whose optimal sets the host needs
    auto run = [&](const std::vector<int32_t> &ie, const std::vector<uint8_t> &ifw, int recode_null, int &W, bool copy_opt) -> int {
        int rc = BNK_OK;
        const int n_inst = (int)ie.size();
        const auto t0 = std::chrono::steady_clock::now();
        BepArgs a;
        std::memset(&a, 0, sizeof(a));
        int32_t flag = 0, mx = 0;
        BEP_CUDA_TRY(cudaMemcpyAsync(d_inst_e.p, ie.data(), sizeof(int32_t) * n_inst, cudaMemcpyHostToDevice, st));
        BEP_CUDA_TRY(cudaMemcpyAsync(d_fwd.p, ifw.data(), n_inst, cudaMemcpyHostToDevice, st));
        BEP_CUDA_TRY(cudaMemsetAsync(d_flag.p, 0, sizeof(int32_t), st));
        a.inst_e = d_inst_e.p; a.inst_fwd = d_fwd.p; a.n_inst = n_inst; a.n_pos = n_pos; a.n_nodes = n; a.n_leaves = n_leaves; a.n_int = n_int;
        a.nxt = d_nxt.p; a.prv = d_prv.p; a.first = d_first.p; a.kids = d_kids.p; a.leaf_of_node = d_leaf_of_node.p;
        a.ent_of_node = d_ent.p; a.parent = d_parent.p; a.recode_null = recode_null; a.max_sym = MAXSYM; a.max_children = t->max_children;
        a.dict = d_dict.p; a.nsym = d_nsym.p; a.A = d_A.p; a.B = d_B.p; a.flag = d_flag.p; a.symmap = d_symmap.p;
        BEP_CUDA_TRY(cudaEventRecord(ev0, st));
        BEP_CUDA_TRY(bep_launch_dict(a, st));
        info[3]++;
        h_nsym.resize(n_inst);
        BEP_CUDA_TRY(cudaMemcpyAsync(h_nsym.data(), d_nsym.p, sizeof(int32_t) * n_inst, cudaMemcpyDeviceToHost, st));
        BEP_CUDA_TRY(cudaMemcpyAsync(&flag, d_flag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        BEP_CUDA_TRY(cudaStreamSynchronize(st));
        if (flag) { rc = fail(BNK_ERR_ARG, "bep: a column has more than %d distinct edge targets", MAXSYM); goto done; }
        for (int x : h_nsym) mx = std::max(mx, x);
        info[2] = std::max(info[2], mx);
        W = mx + 1 <= 32 ? 1 : (mx + 1 <= 64 ? 2 : (mx + 1 <= 128 ? 4 : 8));
        if (by_levels) {
            BEP_CUDA_TRY(bep_launch_sweeps_levels(a, W, d_level_nodes.p, level_off.data(), n_levels, st));
            info[3] += 2 * n_levels;
        } else {
            BEP_CUDA_TRY(bep_launch_sweeps(a, W, st));
            info[3] += 2;
        }
        BEP_CUDA_TRY(cudaEventRecord(ev1, st));
        BEP_CUDA_TRY(pin_dict.ensure(sizeof(int32_t) * (size_t)std::max(mx, 1) * n_inst));
        h_dict = static_cast<const int32_t *>(pin_dict.p);
        BEP_CUDA_TRY(cudaMemcpyAsync(pin_dict.p, d_dict.p, sizeof(int32_t) * (size_t)std::max(mx, 1) * n_inst, cudaMemcpyDeviceToHost, st));
        if (copy_opt && !want_rows.empty()) {
            // only the rows of the ancestors still without a path (r1 copied all n_int rows every round)
            const size_t row_words = (size_t)W * n_inst;
            BEP_CUDA_TRY(d_rows.alloc(want_rows.size()));
            BEP_CUDA_TRY(d_compact.alloc(want_rows.size() * row_words));
            BEP_CUDA_TRY(cudaMemcpyAsync(d_rows.p, want_rows.data(), sizeof(int32_t) * want_rows.size(), cudaMemcpyHostToDevice, st));
            BEP_CUDA_TRY(bep_launch_gather(d_B.p, d_rows.p, (int)want_rows.size(), row_words, d_compact.p, st));
            info[3]++;
            BEP_CUDA_TRY(pin_opt.ensure(sizeof(uint32_t) * want_rows.size() * row_words));
            h_opt = static_cast<const uint32_t *>(pin_opt.p);
            opt_row_of.assign(n_int, -1);
            for (size_t i = 0; i < want_rows.size(); i++) opt_row_of[want_rows[i]] = (int32_t)i;
            BEP_CUDA_TRY(cudaMemcpyAsync(pin_opt.p, d_compact.p, sizeof(uint32_t) * want_rows.size() * row_words, cudaMemcpyDeviceToHost, st));
        } else if (copy_opt) {
            BEP_CUDA_TRY(pin_opt.ensure(sizeof(uint32_t) * (size_t)n_int * W * n_inst));
            h_opt = static_cast<const uint32_t *>(pin_opt.p);
            opt_ent0 = 0;
            opt_row_of.clear();
            BEP_CUDA_TRY(cudaMemcpyAsync(pin_opt.p, d_B.p, sizeof(uint32_t) * (size_t)n_int * W * n_inst, cudaMemcpyDeviceToHost, st));
        }
        BEP_CUDA_TRY(cudaStreamSynchronize(st));
        info[4] += us_since(t0);
        {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) info[7] += (int)(ms * 1000.f);
        }
    done:
        return rc;
    };
    // optimal non-NULL values of ancestor row `ent` in instance k of the last batch
    auto for_each_optimal = [&](int ent, int k, int n_inst, int W, auto &&fn) {
        const int ns = h_nsym[k];
        for (int w = 0; w < W; w++) {
            const size_t orow = opt_row_of.empty() ? (size_t)(ent - opt_ent0) : (size_t)opt_row_of[ent];
            uint32_t bits = h_opt[(orow * W + w) * n_inst + k];
            while (bits) {
                const int b = __builtin_ctz(bits);
                bits &= bits - 1;
                const int sym = w * 32 + b;
                if (sym < ns) fn(h_dict[(size_t)sym * n_inst + k]);  // sym == ns is the NULL symbol
            }
        }
    };
    auto emit_row = [&](int v, const Pog &g) {
        uint8_t *row = present_out + (size_t)v * n_pos;
        for (int c = 0; c < n_pos; c++) row[c] = g.node[c] ? 1 : 0;
    };

    {
        const int n_inst_max = 2 * (n_pos + 1);
        std::vector<int32_t> ie;
        std::vector<uint8_t> ifw;
        int W = 1;
        if (!st) BEP_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        if (!ev0) BEP_CUDA_TRY(cudaEventCreate(&ev0));
        if (!ev1) BEP_CUDA_TRY(cudaEventCreate(&ev1));
        BEP_CUDA_TRY(d_obs.alloc((size_t)n * n_pos));
        BEP_CUDA_TRY(d_leaf_nodes.alloc(n_leaves));
        BEP_CUDA_TRY(d_leaf_of_node.alloc(n));
        BEP_CUDA_TRY(d_ent.alloc(n));
        BEP_CUDA_TRY(d_parent.alloc(n));
        BEP_CUDA_TRY(d_first.alloc(n + 1));
        BEP_CUDA_TRY(d_kids.alloc(t->kids.size()));
        BEP_CUDA_TRY(d_nxt.alloc((size_t)n_leaves * (n_pos + 2)));
        BEP_CUDA_TRY(d_prv.alloc((size_t)n_leaves * (n_pos + 2)));
        BEP_CUDA_TRY(d_inst_e.alloc(n_inst_max));
        BEP_CUDA_TRY(d_fwd.alloc(n_inst_max));
        BEP_CUDA_TRY(d_dict.alloc((size_t)MAXSYM * n_inst_max));
        BEP_CUDA_TRY(d_nsym.alloc(n_inst_max));
        BEP_CUDA_TRY(d_symmap.alloc((size_t)n_inst_max * (n_pos + 2)));
        BEP_CUDA_TRY(d_flag.alloc(1));
        BEP_CUDA_TRY(d_level_nodes.alloc(level_nodes.size()));
        BEP_CUDA_TRY(cudaMemcpyAsync(d_level_nodes.p, level_nodes.data(), sizeof(int32_t) * level_nodes.size(), cudaMemcpyHostToDevice, st));
        BEP_CUDA_TRY(cudaMemcpyAsync(d_obs.p, obs, (size_t)n * n_pos, cudaMemcpyHostToDevice, st));
        BEP_CUDA_TRY(cudaMemcpyAsync(d_leaf_nodes.p, leaf_nodes.data(), sizeof(int32_t) * n_leaves, cudaMemcpyHostToDevice, st));
        BEP_CUDA_TRY(cudaMemcpyAsync(d_leaf_of_node.p, leaf_of_node.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
        BEP_CUDA_TRY(cudaMemcpyAsync(d_ent.p, t->ent_of_node.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
        BEP_CUDA_TRY(cudaMemcpyAsync(d_parent.p, t->parent.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
        BEP_CUDA_TRY(cudaMemcpyAsync(d_first.p, t->first.data(), sizeof(int32_t) * (n + 1), cudaMemcpyHostToDevice, st));
        if (!t->kids.empty())
            BEP_CUDA_TRY(cudaMemcpyAsync(d_kids.p, t->kids.data(), sizeof(int32_t) * t->kids.size(), cudaMemcpyHostToDevice, st));
        mark("allocations + input copies issued");
        BEP_CUDA_TRY(bep_launch_next_prev(d_obs.p, d_leaf_nodes.p, n_leaves, n_pos, d_nxt.p, d_prv.p, st));
        info[3]++;
        // the bit sets of the widest case (8 words) bound the allocation; most alignments need one word
        {
            // allocate for one word first; grow if the dictionary pass asks for more
            // (the first batch decides: the patch batches are subsets of its instances without the NULL symbol)
            BEP_CUDA_TRY(d_A.alloc((size_t)n_int * n_inst_max));
            BEP_CUDA_TRY(d_B.alloc((size_t)n_int * n_inst_max));
        }
        // ---- every (column, direction) instance with RECODE_NULL (Prediction.java:1453-1477) ----
        for (int e = 0; e <= n_pos; e++) { ie.push_back(e); ifw.push_back(1); }       // forward: columns -1 .. nPos - 1
        for (int e = 1; e <= n_pos + 1; e++) { ie.push_back(e); ifw.push_back(0); }   // backward: columns 0 .. nPos
        {
            // dictionary pass once to size the bit sets
            BepArgs a;
            std::memset(&a, 0, sizeof(a));
            int32_t flag = 0, mx = 0;
            const int n_inst = (int)ie.size();
            BEP_CUDA_TRY(cudaMemcpyAsync(d_inst_e.p, ie.data(), sizeof(int32_t) * n_inst, cudaMemcpyHostToDevice, st));
            BEP_CUDA_TRY(cudaMemcpyAsync(d_fwd.p, ifw.data(), n_inst, cudaMemcpyHostToDevice, st));
            BEP_CUDA_TRY(cudaMemsetAsync(d_flag.p, 0, sizeof(int32_t), st));
            a.inst_e = d_inst_e.p; a.inst_fwd = d_fwd.p; a.n_inst = n_inst; a.n_pos = n_pos; a.n_nodes = n; a.n_leaves = n_leaves;
            a.nxt = d_nxt.p; a.prv = d_prv.p; a.max_sym = MAXSYM; a.dict = d_dict.p; a.nsym = d_nsym.p; a.flag = d_flag.p; a.symmap = d_symmap.p;
            BEP_CUDA_TRY(bep_launch_dict(a, st));
            info[3]++;
            h_nsym.resize(n_inst);
            BEP_CUDA_TRY(cudaMemcpyAsync(h_nsym.data(), d_nsym.p, sizeof(int32_t) * n_inst, cudaMemcpyDeviceToHost, st));
            BEP_CUDA_TRY(cudaMemcpyAsync(&flag, d_flag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            BEP_CUDA_TRY(cudaStreamSynchronize(st));
            if (flag) { rc = fail(BNK_ERR_ARG, "bep: a column has more than %d distinct edge targets", MAXSYM); goto done; }
            for (int x : h_nsym) mx = std::max(mx, x);
            const int Wn = mx + 1 <= 32 ? 1 : (mx + 1 <= 64 ? 2 : (mx + 1 <= 128 ? 4 : 8));
            if (Wn > 1) {
                BEP_CUDA_TRY(d_A.alloc((size_t)n_int * Wn * n_inst_max));
                BEP_CUDA_TRY(d_B.alloc((size_t)n_int * Wn * n_inst_max));
            }
        }
        mark("sizing pass done");
        rc = run(ie, ifw, 1, W, false);
        if (rc != BNK_OK) goto done;
        mark("main sweeps done");
        if (device_assembly) {
            const int n_inst = (int)ie.size();
            const auto t0 = std::chrono::steady_clock::now();
            const int n_words = (n_pos + 31) / 32;
            const size_t row_words = (size_t)W * n_inst;
            std::vector<uint8_t> h_contig(n_int);
            std::vector<int32_t> rows;
            BEP_CUDA_TRY(d_A1.alloc((size_t)n_words * n_int));
            BEP_CUDA_TRY(d_Z.alloc((size_t)n_words * n_int));
            BEP_CUDA_TRY(d_contig.alloc(n_int));
            BEP_CUDA_TRY(d_present.alloc((size_t)n * n_pos));
            BEP_CUDA_TRY(cudaEventRecord(ev0, st));
            BEP_CUDA_TRY(bep_launch_reach(d_B.p, d_dict.p, d_nsym.p, info[2], n_int, n_inst, n_pos, W, d_A1.p, d_Z.p, d_contig.p, st));
            BEP_CUDA_TRY(bep_launch_rows(d_obs.p, d_ent.p, d_Z.p, n, n_int, n_pos, d_present.p, st));
            BEP_CUDA_TRY(cudaEventRecord(ev1, st));
            info[3] += 2;
            BEP_CUDA_TRY(cudaMemcpyAsync(h_contig.data(), d_contig.p, n_int, cudaMemcpyDeviceToHost, st));
            BEP_CUDA_TRY(cudaStreamSynchronize(st));
            {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) info[7] += (int)(ms * 1000.f);
            }
            mark("device graph assembly done");
            // ancestors without a start-to-end path: their optimal sets go to the host for the patch rounds
            for (int ent = 0; ent < n_int; ent++) if (!h_contig[ent]) rows.push_back(ent);
            if (!rows.empty()) {
                BEP_CUDA_TRY(d_rows.alloc(rows.size()));
                BEP_CUDA_TRY(d_compact.alloc(rows.size() * row_words));
                BEP_CUDA_TRY(cudaMemcpyAsync(d_rows.p, rows.data(), sizeof(int32_t) * rows.size(), cudaMemcpyHostToDevice, st));
                BEP_CUDA_TRY(bep_launch_gather(d_B.p, d_rows.p, (int)rows.size(), row_words, d_compact.p, st));
                info[3]++;
                BEP_CUDA_TRY(pin_opt.ensure(sizeof(uint32_t) * rows.size() * row_words));
                BEP_CUDA_TRY(cudaMemcpyAsync(pin_opt.p, d_compact.p, sizeof(uint32_t) * rows.size() * row_words, cudaMemcpyDeviceToHost, st));
            }
            // every row of the mask (extants from their residues, ancestors with a path from alive2) in one copy;
            // large masks go through a temporary page-lock of the caller's buffer (a pageable copy runs at ~3 GB/s)
            {
                const size_t mask_bytes = (size_t)n * n_pos;
                const bool pinned = mask_bytes >= ((size_t)8 << 20) && cudaHostRegister(present_out, mask_bytes, cudaHostRegisterDefault) == cudaSuccess;
                if (!pinned) cudaGetLastError();
                cudaError_t ce = cudaMemcpyAsync(present_out, d_present.p, mask_bytes, cudaMemcpyDeviceToHost, st);
                if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
                if (pinned) cudaHostUnregister(present_out);
                BEP_CUDA_TRY(ce);
            }
            mark("mask + rows of ancestors without a path copied");
            if (!rows.empty()) {
                opt_row_of.assign(n_int, -1);
                for (size_t i = 0; i < rows.size(); i++) opt_row_of[rows[i]] = (int32_t)i;
                h_opt = static_cast<const uint32_t *>(pin_opt.p);
                std::mutex mu;
                parallel_slices((int)rows.size(), [&](int slo, int shi) {
                    Pog g;
                    std::vector<Edge3> edges;
                    for (int r = slo; r < shi; r++) {
                        const int ent = rows[r], v = t->node_of_ent[ent];
                        edges.clear();
                        for (int k = 0; k < n_inst; k++) {
                            if (h_nsym[k] < 1) continue;
                            const int i = ie[k] - 1;
                            if (ifw[k]) for_each_optimal(ent, k, n_inst, W, [&](int to) { edges.push_back(Edge3{i, to, 1}); });
                            else for_each_optimal(ent, k, n_inst, W, [&](int from) { edges.push_back(Edge3{from, i, 2}); });
                        }
                        merge_edges(edges);
                        g.from_edges(n_pos, edges);
                        if (!g.is_contiguous()) {
                            std::set<int> pr = g.protruding();
                            std::lock_guard<std::mutex> lk(mu);
                            crippled[v] = std::move(pr);
                            crippled_edges[v] = edges;
                        } else {   // cannot happen (the device found no path); kept so that a disagreement is harmless
                            g.nibble();
                            emit_row(v, g);
                        }
                    }
                });
                opt_row_of.clear();
            }
            info[5] += us_since(t0);
        } else {
            const int n_inst = (int)ie.size();
            const auto t0 = std::chrono::steady_clock::now();
            std::mutex mu;
            // the optimal sets come back in chunks of ancestor rows through two page-locked buffers: the copy of
            // chunk c + 1 runs while the host's threads assemble the graphs of chunk c
            const size_t row_words = (size_t)W * n_inst;
            const int chunk_rows = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_int, ((size_t)96 << 20) / (row_words * 4)));
            const int n_chunks = (n_int + chunk_rows - 1) / chunk_rows;
            BEP_CUDA_TRY(pin_opt.ensure(2 * (size_t)chunk_rows * row_words * 4));
            uint32_t *buf[2] = {static_cast<uint32_t *>(pin_opt.p), static_cast<uint32_t *>(pin_opt.p) + (size_t)chunk_rows * row_words};
            cudaEvent_t cev[2] = {nullptr, nullptr};
            BEP_CUDA_TRY(cudaEventCreate(&cev[0]));
            BEP_CUDA_TRY(cudaEventCreate(&cev[1]));
            auto issue = [&](int c) -> cudaError_t {
                const int lo = c * chunk_rows, hi = std::min(n_int, lo + chunk_rows);
                cudaError_t e = cudaMemcpyAsync(buf[c & 1], d_B.p + (size_t)lo * row_words, (size_t)(hi - lo) * row_words * 4, cudaMemcpyDeviceToHost, st);
                return e != cudaSuccess ? e : cudaEventRecord(cev[c & 1], st);
            };
            cudaError_t ce = issue(0);
            for (int c = 0; c < n_chunks && ce == cudaSuccess; c++) {
                if (c + 1 < n_chunks) ce = issue(c + 1);
                if (ce == cudaSuccess) ce = cudaEventSynchronize(cev[c & 1]);
                if (ce != cudaSuccess) break;
                const int lo = c * chunk_rows, hi = std::min(n_int, lo + chunk_rows);
                h_opt = buf[c & 1];
                opt_ent0 = lo;
                parallel_slices(hi - lo, [&](int slo, int shi) {
                    Pog g;
                    std::vector<Edge3> edges;
                    for (int ent = lo + slo; ent < lo + shi; ent++) {
                        const int v = t->node_of_ent[ent];
                        edges.clear();
                        for (int k = 0; k < n_inst; k++) {
                            if (h_nsym[k] < 1) continue;  // nothing to infer: the reference makes no Parsimony object (:1466-1468)
                            const int i = ie[k] - 1;
                            if (ifw[k]) for_each_optimal(ent, k, n_inst, W, [&](int to) { edges.push_back(Edge3{i, to, 1}); });
                            else for_each_optimal(ent, k, n_inst, W, [&](int from) { edges.push_back(Edge3{from, i, 2}); });
                        }
                        merge_edges(edges);
                        g.from_edges(n_pos, edges);
                        if (!g.is_contiguous()) {
                            std::set<int> pr = g.protruding();
                            std::lock_guard<std::mutex> lk(mu);
                            crippled[v] = std::move(pr);
                            crippled_edges[v] = edges;
                        } else {
                            g.nibble();  // GRASP.NIBBLE = true
                            emit_row(v, g);
                            if (pogs) g.export_to(pogs, ent, edges);
                        }
                    }
                });
            }
            cudaStreamSynchronize(st);
            cudaEventDestroy(cev[0]);
            cudaEventDestroy(cev[1]);
            BEP_CUDA_TRY(ce);
            info[5] += us_since(t0);
        }
        info[0] = (int)crippled.size();
        mark("first-pass graph assembly done");
        // ---- patchAncestorsWithBEP: re-infer the protruding columns without the NULL symbol (:1551-1588) ----
        for (;;) {
            std::set<int> columns;
            for (const auto &kv : crippled) columns.insert(kv.second.begin(), kv.second.end());
            if (columns.empty()) break;
            const std::vector<int> cols(columns.begin(), columns.end());
            ie.clear(); ifw.clear();
            for (int idx : cols) { ie.push_back(std::abs(idx) + 1); ifw.push_back(idx > 0 ? 1 : 0); }
            want_rows.clear();
            for (const auto &kv : crippled) want_rows.push_back(t->ent_of_node[kv.first]);
            rc = run(ie, ifw, 0, W, true);
            if (rc != BNK_OK) goto done;
            info[1]++;
            const auto t0 = std::chrono::steady_clock::now();
            bool changed = false;
            const int n_inst = (int)ie.size();
            std::map<int, int> inst_of;  // protruding index -> instance of this round
            for (int k = 0; k < n_inst; k++) inst_of[cols[k]] = k;
            std::vector<int> fixme;
            for (const auto &kv : crippled) fixme.push_back(kv.first);
            std::vector<char> fixed(fixme.size(), 0), grew(fixme.size(), 0);
            std::vector<std::set<int>> next_pr(fixme.size());
            parallel_slices((int)fixme.size(), [&](int lo, int hi) {
                Pog g;
                for (int f = lo; f < hi; f++) {
                    const int v = fixme[f];
                    std::vector<Edge3> &edges = crippled_edges.find(v)->second;  // maps are only read here
                    g.from_edges(n_pos, edges);
                    const int ent = t->ent_of_node[v];
                    for (int idx : crippled.find(v)->second) {  // ascending, as cols_ordered
                        const int k = inst_of.find(idx)->second, col = std::abs(idx);
                        for_each_optimal(ent, k, n_inst, W, [&](int inferred) {
                            if (!g.is_node(inferred) && inferred != n_pos && inferred != -1) { g.add_node(inferred); grew[f] = 1; }
                            // pog.addEdge(.., new BidirEdge(forward, !forward)) REPLACES the edge object of an existing edge
                            // (IdxEdgeGraph.addEdge, src/dat/pog/IdxEdgeGraph.java:147-153): the flags are overwritten
                            const Edge3 ed = idx > 0 ? Edge3{col, inferred, 1} : Edge3{inferred, col, 2};
                            bool known = false;
                            for (Edge3 &old : edges) if (old.a == ed.a && old.b == ed.b) { old.f = ed.f; known = true; }
                            if (g.add_edge(ed.a, ed.b)) { if (!known) edges.push_back(ed); grew[f] = 1; }
                        });
                    }
                    if (!g.is_contiguous()) {
                        next_pr[f] = g.protruding();
                    } else {
                        g.nibble();
                        emit_row(v, g);
                        if (pogs) g.export_to(pogs, ent, edges);
                        fixed[f] = 1;
                    }
                }
            });
            for (size_t f = 0; f < fixme.size(); f++) {
                changed |= grew[f] != 0;
                if (fixed[f]) { crippled.erase(fixme[f]); crippled_edges.erase(fixme[f]); }
                else crippled[fixme[f]] = std::move(next_pr[f]);
            }
            info[5] += us_since(t0);
            // the reference loops `while (columns.size() > 0)` and would spin if a round adds nothing; stop instead
            if (!changed) break;
        }
        for (const auto &kv : crippled_edges) {  // still without a path: nodes as they stand
            pog.from_edges(n_pos, kv.second);
            emit_row(kv.first, pog);
            if (pogs) pog.export_to(pogs, t->ent_of_node[kv.first], kv.second);
        }
        mark("patch rounds done");
        info[6] = us_since(t_begin);
        if (info_or_null) std::memcpy(info_or_null, info, sizeof(info));
    }
done:
    mark("before cleanup");
    if (st) cudaStreamSynchronize(st);
    S.trim();
    return rc;
}


extern "C" int bnk_bep_presence(const bnk_tree *t, int32_t n_cols, const uint8_t *obs, int device, uint8_t *present_out,
                                int32_t *info_or_null)
{
    return bep_run(t, n_cols, obs, device, present_out, info_or_null, nullptr);
}

extern "C" int bnk_bep_infer(const bnk_tree *t, int32_t n_cols, const uint8_t *obs, int device, uint8_t *present_out,
                             int32_t *info_or_null, bnk_pogs **out)
{
    if (!out) return fail(BNK_ERR_ARG, "bep: null argument");
    if (!t) return fail(BNK_ERR_ARG, "bep: bad argument");
    bnk_pogs *p = new (std::nothrow) bnk_pogs();
    if (!p) return fail(BNK_ERR_NOMEM, "bep: out of host memory");
    p->n_pos = n_cols;
    p->ent_of_node = t->ent_of_node;
    p->from.assign(t->n_int, {}); p->to.assign(t->n_int, {}); p->flags.assign(t->n_int, {});
    const int rc = bep_run(t, n_cols, obs, device, present_out, info_or_null, p);
    if (rc != BNK_OK) { delete p; return rc; }
    *out = p;
    return BNK_OK;
}

extern "C" void bnk_pogs_destroy(bnk_pogs *p) { delete p; }

extern "C" int bnk_pogs_n_edges(const bnk_pogs *p, int32_t node)
{
    if (!p || node < 0 || node >= (int32_t)p->ent_of_node.size() || p->ent_of_node[node] < 0)
        return fail(BNK_ERR_QUERY, "pogs: node %d is not an ancestor", node);
    return (int)p->from[p->ent_of_node[node]].size();
}

extern "C" int bnk_pogs_edges(const bnk_pogs *p, int32_t node, int32_t *from, int32_t *to, uint8_t *flags)
{
    const int n = bnk_pogs_n_edges(p, node);
    if (n < 0) return n;
    const int ent = p->ent_of_node[node];
    if (from) std::memcpy(from, p->from[ent].data(), sizeof(int32_t) * n);
    if (to) std::memcpy(to, p->to[ent].data(), sizeof(int32_t) * n);
    if (flags) std::memcpy(flags, p->flags[ent].data(), n);
    return n;
}
