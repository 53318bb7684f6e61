// device_common.cuh -- types shared by the CUDA translation units of libbnkgpu.
#pragma once
#include "internal.h"

#include <cuda_runtime.h>
#include <math_constants.h>

namespace bnk {

struct ModelDev {  // device pointers
    const double *evec, *ievec, *eval;  // K*K, K*K, K
    const double *prior, *logprior;     // K each
    int K;
};

// Per (category, node) tables, index ((cat * n_nodes + node) * K + a) * K + b.
//   P [a][b] = P(child = b | parent = a)      PT[a][b] = P [b][a]
//   L = log P (fdlibm log, log 0 = -inf)      LT likewise
struct TableSet {
    double *P, *PT, *L, *LT;
};

// A tile = a run of consecutive columns (in category-sorted order) owned by one CTA.
struct TileDesc {
    int32_t col0, ncols;  // sorted-space column range
    int32_t cat0, ncat;   // chunk-local category range covered by those columns
};

enum : uint8_t { KIND_NONE = 0, KIND_MSG = 1, KIND_ROOT = 2, KIND_OBS = 3 };

struct WalkArgs {
    // static program
    const Step *up;
    const Step *down;
    const ChildRef *up_child, *down_child;
    int32_t n_steps;
    const TileDesc *tiles;
    int32_t ncg;       // column groups per CTA (threads = ncg * K, rounded up to a warp)
    int32_t ncat_max;  // largest TileDesc::ncat
    // inputs in the CALLER's (original) column order, [node][ncols]; kernels read them through orig_col
    const uint8_t *obs, *present;  // present == nullptr: everything present
    int32_t ncols, n_nodes;
    const int32_t *orig_col;  // sorted -> original column index
    const int32_t *col_cat;   // sorted column -> chunk-local category
    // tables of the current category chunk
    const double *T_row;  // matvec rows: L (joint up), P (sum up), PT (down)
    const double *T_col;  // lookups [obs][x]: LT (joint), PT (sum up / down)
    const double *T_aux;  // warp-tiled down-sweep (sweeps_warp.cu): P, read as rows [p][x .. x + 3]
    const double *prior;  // log prior (joint) or prior (sum)
    // message stack: slots < smem_slots live in shared memory, the rest in spill[(slot - smem_slots)][ncols][K]
    int32_t smem_slots;
    double *spill;
    // joint intermediates, sorted column space
    uint8_t *ptr;   // [n_int][ncols][K] traceback pointers
    uint8_t *kind;  // [n_int][ncols]
    // sum-product intermediates, sorted column space
    double *msg;     // [n_int][ncols][K] scaled up-messages (nullptr: not stored)
    uint8_t *kindm;  // [n_int][ncols] (GENERAL only)
    // sum-product outputs, original column space
    double *out_logl;       // [ncols]
    double *out_post;       // [n_query][ncols][K]
    const int32_t *qrow;    // [n_int] -> row of out_post or -1
    // hybrid schedule (lean kernels only): children handled by the wide level kernels
    double *tmsg;               // from-above vectors handed to wide children, [ent][K][ncols]
    const int32_t *esum_wide;   // [ncols] rescale exponents accumulated by the wide sum up-sweep (or nullptr)
    const double *logl_wide;    // [ncols] log-likelihood terms the general wide kernels accumulated (or nullptr)
    // masked warp-tiled walker (sweeps_warp.cu): case byte of every (walker node, sorted column), [n_int][ncols]:
    // kind | child 0 present << 2 | child 1 present << 3 (launch_walk_case); nullptr selects the unmasked kernels
    const uint8_t *caseb;
    // general joint up-sweep, nodes with >= 32 children: queue of the pairwise additions, one region per (thread, column)
    double *qscratch;
    int32_t qstride;
};

// a height-2 node with the height-1 children the fused down-sweep kernel (wide.cu: down_wide2_kernel) evaluates in
// the same thread; c[e].node = -1: child e is childless (or absent) and nothing is fused for it
struct Wide2Node {
    WideNode v;
    WideNode c[2];
    int32_t crow[2];  // row of c[e] in the height-1 level (cherry-direct: its message lives in the record tables)
};

// arguments of the level kernels (wide.cu); all column indices are in sorted space
struct WideArgs {
    const WideNode *nodes;   // the level (or a slice of it): gridDim.y entries
    const TileDesc *tiles;   // wide column tiles: one category each, <= 128 columns
    int32_t n_tiles, tiles_per_cta;  // a CTA walks tiles_per_cta consecutive tiles (tables reloaded only when the category changes)
    const uint8_t *obs;
    int32_t ncols, n_nodes;
    const int32_t *orig_col;
    const double *T_row, *T_col;
    double *msg;             // wide nodes: [ent][K][ncols]
    uint8_t *ptr;            // wide nodes: [ent][K][ncols]
    int16_t *escale;         // [wide row][ncols] rescale exponent of each (node, column)
    int32_t wide_row0;       // wide row of nodes[0]
    int32_t cherry;          // the level is height 1: every child is childless
    const TileDesc *cgroups; // cherry.cu: column groups of one category, <= cherry_group_cols() columns (nullptr: not used)
    int32_t n_cgroups;
    uint8_t *cscratch;       // cherry.cu: result records of the distinct pairs, [level node][crecs_per_node] records
    int32_t crecs_per_node;  //   (the record offset of a group inside a node's region is cgroups[].ncat)
    int16_t *cslots;         // cherry.cu: record index of every column, [level node][ncols]
    const TileDesc *ctiles;  // cherry.cu expansion tiles: col0 = 16-aligned start, cat0 = group
    int32_t n_ctiles;
    double *tmsg;
    double *out_post;
    const int32_t *qrow;
    uint8_t *state_s;        // [node][ncols]
    int32_t col_lo, col_hi;  // traceback: columns of the chunk being processed
    // general level kernels (wide_general.cu)
    const uint8_t *present;  // original column order, or nullptr
    const double *prior;     // log prior (joint) / prior (sum)
    uint8_t *kind, *kindm;   // [ent][ncols] node kinds (joint / sum)
    double *logl_acc;        // [ncols] += log of component-root totals and scalar factors (atomic)
    const Wide2Node *fuse2;  // down_wide2_kernel: the height-2 level (or a slice of it) with its fused children
    // cherry-direct (wide_common.cuh): nullptr = every height-1 message was expanded into a.msg
    const int32_t *cherry_row;  // [n_int] ent -> row in the height-1 level of a node whose message stays in the record tables, else -1
    const int32_t *tile_group;  // [n_tiles] first cherry group overlapping a.tiles[i]
    const int32_t *rest_rows;   // cherry_expand_kernel on a subset: row in the height-1 level of a.nodes[i]
};
cudaError_t launch_down_wide2(int K, const WideArgs &a, int n_tiles, int n_nodes, cudaStream_t st);
cudaError_t launch_up_wide_gen(int K, int mode, const WideArgs &a, int n_tiles, int n_nodes, cudaStream_t st);
cudaError_t launch_down_wide_gen(int K, const WideArgs &a, int n_tiles, int n_nodes, cudaStream_t st);
cudaError_t launch_traceback_wide_gen(int K, const WideArgs &a, int n_nodes, cudaStream_t st);
cudaError_t launch_up_wide(int K, int mode, const WideArgs &a, int n_tiles, int n_nodes, cudaStream_t st);
cudaError_t launch_down_wide(int K, const WideArgs &a, int n_tiles, int n_nodes, cudaStream_t st);
cudaError_t launch_traceback_wide(int K, const WideArgs &a, int n_nodes, cudaStream_t st);
cudaError_t launch_sum_escale(const int16_t *escale, int n_rows, int ncols, int32_t *esum, cudaStream_t st);
cudaError_t launch_scatter_rows(const uint8_t *in, uint8_t *out, const int32_t *nodes, int n_rows,
                                const int32_t *orig_col, int ncols, cudaStream_t st);
int wide_tile_cols();
cudaError_t launch_up_cherry(int K, int mode, const WideArgs &a, int n_nodes, cudaStream_t st);
cudaError_t launch_cherry_pairs(int K, int mode, const WideArgs &a, int n_nodes, cudaStream_t st);
cudaError_t launch_cherry_expand(int K, int mode, const WideArgs &a, int n_nodes, cudaStream_t st);
cudaError_t launch_traceback_cherry(int K, const WideArgs &a, int n_nodes, cudaStream_t st);
int cherry_group_cols();
size_t cherry_record_bytes(int K);
int cherry_group_capacity(int K, int ncols);

void launch_build_tables(int K, const ModelDev &md, const double *times, const int32_t *skip,
                         const double *rates, int n_nodes, int n_cat, const TableSet &ts, cudaStream_t st);
void launch_log_prior(const double *prior, double *logprior, int K, cudaStream_t st);

struct WalkLaunch {
    int K, cpt, threads, n_tiles;
    bool general;
    size_t smem_bytes;
    cudaStream_t stream;
};
size_t walk_smem_bytes(int K, int cpt, int ncg, int ncat_max, int slots);
int walk_max_smem_optin(int device);
cudaError_t launch_joint_up(const WalkLaunch &wl, const WalkArgs &a);
cudaError_t launch_sum_up(const WalkLaunch &wl, const WalkArgs &a);
cudaError_t launch_sum_down(const WalkLaunch &wl, const WalkArgs &a);
// lean variants (sweeps_fast.cu): no mask, evidence on childless nodes only, <= 2 children per
// node, one rate category per tile, whole message stack in shared memory
size_t fast_smem_bytes(int K, int cpt, int ncg, int slots);
cudaError_t launch_joint_up_fast(const WalkLaunch &wl, const WalkArgs &a);
cudaError_t launch_sum_up_fast(const WalkLaunch &wl, const WalkArgs &a);
cudaError_t launch_sum_down_fast(const WalkLaunch &wl, const WalkArgs &a);

// warp-tiled walker (sweeps_warp.cu): K = 20, same eligibility as the lean kernels; a.tiles = tiles of one
// category and at most W * warp_walk_cols_per_warp(C) (<= 64) columns; W consumer warps (+ 1 producer warp),
// C columns per lane, ns ring stages
size_t warp_walk_smem_bytes(int W, int C, int ns, int slots, bool down, bool masked = false);
cudaError_t launch_walk_case(const Step *prog, int n_steps, const uint8_t *present, const int32_t *orig_col, int ncols,
                             uint8_t *out, cudaStream_t st);
int warp_walk_cols_per_warp(int C);
cudaError_t launch_joint_up_warp(const WalkArgs &a, int n_tiles, int W, int C, int ns, cudaStream_t st);
cudaError_t launch_sum_up_warp(const WalkArgs &a, int n_tiles, int W, int C, int ns, cudaStream_t st);
cudaError_t launch_sum_down_warp(const WalkArgs &a, int n_tiles, int W, int C, int ns, cudaStream_t st);

// sum-product sweeps of the warp-tiled walker on the FP64 tensor-core path (DMMA.8x8x4): 8 columns per warp,
// a.tiles = tiles of one category and at most 8 W columns
size_t mma_walk_smem_bytes(int W, int ns, int slots, bool down);
int mma_walk_cols_per_warp();
cudaError_t launch_sum_up_mma(const WalkArgs &a, int n_tiles, int W, int ns, cudaStream_t st);

struct TracebackArgs {
    const Step *down;
    int32_t n_steps, ncols, n_nodes, K;
    // walk traceback: everything in SORTED column space
    const uint8_t *ptr, *kind;      // kind == nullptr: root steps are KIND_ROOT, all others KIND_MSG
    int32_t kind_obs;               // 0: no kind byte is KIND_OBS (evidence on childless nodes only)
    const int32_t *orig_col;        // sorted -> original column (KIND_OBS pass-through reads obs)
    uint8_t *state_s;               // [node][ncols] sorted space
    int32_t col_lo, col_hi;         // sorted columns of the chunk being processed
    int32_t tb_slots;
    // leaf pass: ORIGINAL column space
    const uint8_t *obs, *present;   // [node][ncols]
    uint8_t *out_state;             // [node][ncols]
    // childless nodes
    const int32_t *leaf_nodes, *leaf_parent;
    int32_t n_leaves;
    const double *L;                // [cat][node][K*K] for the rare unobserved-leaf case
    const int32_t *cat_of_col;      // original column -> global category
    int32_t cat_lo, cat_hi;         // categories of the chunk being processed; other columns are skipped
};
cudaError_t launch_traceback(const TracebackArgs &a, cudaStream_t st);
cudaError_t launch_leaf_states(const TracebackArgs &a, cudaStream_t st);

struct LeafPostArgs {
    int32_t node, parent, ncols, n_nodes, K;
    const uint8_t *obs, *present;   // original column order
    const double *P;                // [cat][node][K*K] tables of the chunk being processed
    const int32_t *cat_of_col;      // original column -> global category
    int32_t cat_lo, cat_hi;
    const double *parent_row;       // [ncols][K] posterior of the parent (NaN rows where it has none)
    double *out_row;                // [ncols][K]
};
cudaError_t launch_leaf_post(const LeafPostArgs &a, cudaStream_t st);

cudaError_t launch_gather_cols(const uint8_t *in, uint8_t *out, const int32_t *orig_col, int n_rows,
                               int ncols, cudaStream_t st);
// flag[0] |= 1 if a node with children holds an observation, |= 2 if a byte is neither BNK_NONE nor < K
cudaError_t launch_scan_obs(const uint8_t *obs, const uint8_t *has_child, int n_nodes, int ncols, int K,
                            int32_t *flag, cudaStream_t st);
cudaError_t launch_fill_f64(double *p, size_t n, double v, cudaStream_t st);
cudaError_t launch_sum_f64(const double *in, int n, double *out, cudaStream_t st);

}  // namespace bnk
