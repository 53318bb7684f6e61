// workspace.cuh -- the workspace object behind the C ABI (include/bnkgpu.h): device buffers, rate categories,
// column tiles.  Shared by api.cu (single-device entry points) and multi.cu (column-chunk pipeline, multi-GPU).
#pragma once
#include "device_common.cuh"

#include <string>
#include <utility>
#include <vector>

namespace bnk {

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(e_ == cudaErrorMemoryAllocation ? BNK_ERR_NOMEM : BNK_ERR_CUDA, "%s: %s (%s:%d)", \
                        #expr, cudaGetErrorString(e_), __FILE__, __LINE__);                              \
    } while (0)

// Every device buffer of a workspace registers itself with the workspace's BufRegistry: one place that
// accounts the bytes held (bnk_ws_device_bytes) and releases everything (ws_free).
struct DevBufBase {
    virtual void release() = 0;
    virtual ~DevBufBase() {}
};
struct BufRegistry {
    std::vector<DevBufBase *> all;
    int64_t bytes = 0;
};

template <typename T>
struct DevBuf : DevBufBase {
    T *p = nullptr;
    size_t cap = 0;  // elements
    BufRegistry *reg = nullptr;
    DevBuf() {}
    explicit DevBuf(BufRegistry *r) : reg(r) { if (r) r->all.push_back(this); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = n;
        if (reg) reg->bytes += (int64_t)(n * sizeof(T));
        return cudaSuccess;
    }
    void release() override
    {
        if (p) { cudaFree(p); if (reg) reg->bytes -= (int64_t)(cap * sizeof(T)); }
        p = nullptr; cap = 0;
    }
};

// Entry points select the workspace's device and put the caller's current device back on exit (a host
// process that shares the thread with torch or with other workspaces must not find itself re-targeted).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != device) err = cudaSetDevice(device);
    }
    ~DeviceGuard()
    {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};
#define ENTER_DEVICE(dev)      \
    DeviceGuard dev_guard_(dev); \
    CUDA_TRY(dev_guard_.err)

struct Chunk {
    int cat_lo, cat_hi;    // global category range
    int tile_lo, tile_hi;  // tiles of this chunk
    int ncat_max;          // largest category span of one of its tiles
};

}  // namespace bnk

using namespace bnk;  // internal header: only libbnkgpu translation units include it

struct bnk_ws {
    BufRegistry reg;  // declared first: the buffers below register themselves here
    bnk_model model;
    const bnk_tree *tree = nullptr;
    int K = 0, n = 0, n_int = 0, max_cols = 0, device = 0, sm_count = 148, smem_optin = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t launches = 0;
    std::vector<bnk_ws *> subs;  // column-chunk workspaces of the pipelined host-buffer path (owned)
    cudaStream_t copy_stream = nullptr;          // result copies of the pipelined path: one queue, so that a chunk's D2H never
    cudaEvent_t chunk_done[2] = {}, copy_done[2] = {};  // shares PCIe with its neighbour's and kernels restart as soon as a buffer is free
    int32_t pipeline_cols = 0;   // bnk_ws_set_pipeline: 0 automatic, < 0 off, > 0 columns per chunk
    int32_t sub_cols = 0;        // max_cols the sub-workspaces were created with
    uint64_t dist_version = 0, sub_dist_version = 0;  // bnk_set_distances calls seen / forwarded to the subs
    bool is_sub = false;
    // bnk_joint_marginal_dev: the joint pass runs in a twin workspace on its own stream, beside this workspace's
    // marginal pass (fork / join through events); phases 1, 2 of bnk_ws_phase_ms then come from the twin
    bnk_ws *twin = nullptr;
    cudaEvent_t twin_fork = nullptr, twin_join = nullptr;
    uint64_t twin_dist_version = 0;
    bool twin_phases = false;
    int cfg_tile_cols = 0, cfg_cpt = 0;
    bool force_generic = false;
    int wide_tpc_max = 8;  // r1 sweep on C3a: 4-8 tiles per CTA best
    size_t table_budget = (size_t)12 << 30;
    // static device data
    DevBuf<double> d_model{&reg};  // evec | ievec | eval | prior | logprior
    DevBuf<Step> d_up{&reg}, d_down{&reg}, d_up_n{&reg}, d_down_n{&reg};             // full / narrow walk programs
    DevBuf<ChildRef> d_up_child{&reg}, d_down_child{&reg}, d_up_child_n{&reg}, d_down_child_n{&reg};
    DevBuf<WideNode> d_wide{&reg};                                  // wide levels, concatenated (height 1 first)
    // fused down-sweep (wide.cu: down_wide2_kernel): the wide nodes above height 1 with their height-1 children,
    // and the height-1 nodes whose parent is NOT a wide node (they keep the plain level kernel)
    DevBuf<Wide2Node> d_fuse2{&reg};
    DevBuf<WideNode> d_wide_rest{&reg};
    int n_fuse2 = 0, n_wide_rest = 0;
    DevBuf<int32_t> d_rest_rows{&reg}, d_cherry_row{&reg}, d_wtile_group{&reg};  // cherry-direct (wide_common.cuh)
    std::vector<int32_t> wtile_group;
    bool no_cherry_direct = false;                            // BNK_CHERRY_DIRECT=0 (A/B, tests)
    std::vector<int32_t> fuse_off;                            // [wide_h + 1] into d_fuse2, per height level (index 1 = height 2)
    bool no_fuse2 = false;                                    // BNK_NO_FUSE2=1 (A/B, tests)
    std::vector<int32_t> level_off;                           // [wide_h + 1] into d_wide
    bool no_wide = false;                                     // bnk_ws_configure: walker only (tests, profiling)
    DevBuf<int32_t> d_parent{&reg}, d_node_of_ent{&reg}, d_leaf_nodes{&reg}, d_leaf_parent{&reg};
    DevBuf<uint8_t> d_has_child{&reg};                        // [n] 1 = the node has children (an ancestor)
    int evidence_mode = BNK_EVIDENCE_AUTO;                    // bnk_ws_set_evidence_mode
    bool masked_only = false;                                 // last prepare_inputs: presence mask, evidence on childless nodes only
    bool retile = false;                                      // bnk_ws_configure changed the geometry: re-tile the current rates
    DevBuf<double> d_dist{&reg};
    std::vector<double> dist;
    int n_leaves = 0;
    // rates / categories / tiles
    int ncols = -1;  // columns of the current rate assignment
    bool identity_perm = true;
    std::vector<double> cat_rate, last_rates;
    bool rates_null = true;
    std::vector<int32_t> orig_col, col_cat_global, cat_of_col;
    std::vector<TileDesc> tiles, wtiles;   // walker tiles; wide tiles (one category, <= 128 columns)
    std::vector<Chunk> chunks;
    std::vector<TileDesc> cgroups;         // cherry groups (one category, <= cherry_group_cols() columns)
    std::vector<std::pair<int, int>> chunk_cgroups;
    DevBuf<TileDesc> d_cgroups{&reg};
    std::vector<int> chunk_crecs;          // per chunk: result records per height-1 node (sum of the groups' capacities)
    DevBuf<uint8_t> d_cscratch{&reg};
    DevBuf<int16_t> d_cslots{&reg};
    std::vector<TileDesc> ctiles;          // expansion tiles of the cherry groups (16-column aligned starts)
    std::vector<std::pair<int, int>> chunk_ctiles;
    DevBuf<TileDesc> d_ctiles{&reg};
    // warp-tiled walker (sweeps_warp.cu): tiles of one category, <= 6 * xw columns
    std::vector<TileDesc> xtiles, mtiles;  // mtiles: DMMA sum-product up-sweep (<= 8 * mw columns)
    std::vector<std::pair<int, int>> chunk_xtiles, chunk_mtiles;
    DevBuf<TileDesc> d_xtiles{&reg}, d_mtiles{&reg};
    int mw = 0;                            // consumer warps per CTA of the DMMA variants
    int xw = 0, xc = 2;                    // consumer warps per CTA, columns per lane
    int walker_mode = 0;                   // BNK_WALKER: 0 auto (= warp-tiled + DMMA), 1 lean CTA walker,
                                           // 2 warp-tiled with the scalar sum-product kernels (no DMMA), 3 warp-tiled (r1 default)
    int cfg_xw = 0, cfg_xns = 0, cfg_xc = 0;  // BNK_WARP_W / BNK_WARP_NS / BNK_WARP_C (tuning knobs, tests)
    bool no_cherry = false;                // BNK_NO_CHERRY=1: height-1 level through the plain level kernel (A/B, tests)
    std::vector<std::pair<int, int>> chunk_wtiles, chunk_cols;  // per chunk: wide tile range, sorted column range
    int tile_cols = 0, cpt = 1, ncg = 0, threads = 0;
    DevBuf<double> d_cat_rate{&reg};
    DevBuf<int32_t> d_orig_col{&reg}, d_col_cat{&reg}, d_cat_of_col{&reg}, d_qrow{&reg}, d_flag{&reg};
    DevBuf<TileDesc> d_tiles{&reg}, d_wtiles{&reg};
    DevBuf<int16_t> d_escale{&reg};
    DevBuf<int32_t> d_esum_wide{&reg};
    bool tables_log_valid = false, tables_lin_valid = false;  // only meaningful with a single chunk
    // buffers
    DevBuf<double> d_P{&reg}, d_PT{&reg}, d_L{&reg}, d_LT{&reg}, d_spill{&reg}, d_msg{&reg}, d_tmsg{&reg}, d_logl{&reg}, d_logl_wide{&reg}, d_post{&reg}, d_post_x{&reg}, d_sum{&reg}, d_qscratch{&reg};
    DevBuf<uint8_t> d_obs_in{&reg}, d_present_in{&reg}, d_obs_s{&reg}, d_present_s{&reg}, d_ptr{&reg}, d_kind{&reg}, d_kindm{&reg}, d_state{&reg}, d_state_s{&reg};
    cudaEvent_t ev[5][2] = {};
    bool ev_valid[5] = {false, false, false, false, false};
    ModelDev md() const
    {
        const double *b = d_model.p;
        return ModelDev{b, b + K * K, b + 2 * K * K, b + 2 * K * K + K, b + 2 * K * K + 2 * K, K};
    }
};


#define TRY(expr)                       \
    do {                                \
        int rc_ = (expr);               \
        if (rc_ != BNK_OK) return rc_;  \
    } while (0)

namespace bnk {
// api.cu
int ws_run_sum(bnk_ws *ws, int32_t n_cols, const uint8_t *d_obs, const uint8_t *d_present, const int32_t *query_nodes,
               int32_t n_query, double *d_out_post, double *d_out_logl, double *d_out_sum);
// multi.cu: the host-buffer entry points run column chunks through two sub-workspaces on their own streams so
// that the result copy of one chunk overlaps the kernels of the next (op: 0 joint, 1 marginal, 2 log-likelihood)
struct HostCall {
    int op;
    int32_t n_cols_total;            // row stride of the host arrays, in columns
    int32_t col0, ncols;             // the column range this call covers
    const uint8_t *obs, *present;    // [n][n_cols_total]
    const double *rates;             // [n_cols_total] or nullptr
    uint8_t *out_state;              // [n][n_cols_total]
    const int32_t *query_nodes;
    int32_t n_query;
    double *out_post;                // [n_query][n_cols_total][K]
    double *out_logl;                // [n_cols_total]
    double *out_sum;                 // += sum over the range (host double)
};
bool ws_wants_pipeline(const bnk_ws *ws, int32_t n_cols);
int ws_run_host_range(bnk_ws *ws, const HostCall &hc);
void ws_free(bnk_ws *ws);
}  // namespace bnk
