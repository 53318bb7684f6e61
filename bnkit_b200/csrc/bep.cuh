// bep.cuh -- shared by bep.cu (kernels) and bep_api.cu (host side of bnk_bep_presence)
#pragma once
#include "device_common.cuh"

#include <climits>

namespace bnk {

constexpr int BEP_NONE = INT_MIN;  // "null" leaf value: the sequence has no residue in this column

struct BepArgs {
    const int32_t *inst_e;    // [n_inst] extended column index of the instance
    const uint8_t *inst_fwd;  // [n_inst] 1 = forward
    int32_t n_inst, n_pos, n_nodes, n_leaves, n_int;
    const int32_t *nxt, *prv;            // [leaf][n_pos + 2]
    const int32_t *first, *kids;         // children CSR
    const int32_t *leaf_of_node;         // node -> leaf row or -1
    const int32_t *ent_of_node;          // node -> internal row or -1
    const int32_t *parent;
    int32_t max_children;                // largest number of children of a node (selects the counter width)
    int32_t recode_null;                 // GRASP.RECODE_NULL: childless nodes without a value carry a NULL symbol
    int32_t max_sym;                     // dictionary capacity (32 * W - 1: one bit is kept for NULL)
    uint8_t *symmap;                     // [n_inst][n_pos + 2] value + 1 -> symbol + 1 (0: no leaf has the value)
    int32_t *dict;                       // [max_sym][n_inst] distinct leaf values of the instance
    int32_t *nsym;                       // [n_inst] number of them
    uint32_t *A, *B;                     // [n_int][W][n_inst]; B is overwritten by the optimal sets going down
    int32_t *flag;                       // |= 1: an instance has more than max_sym distinct values
};

cudaError_t bep_launch_next_prev(const uint8_t *obs, const int32_t *leaf_nodes, int n_leaves, int n_pos, int32_t *nxt, int32_t *prv,
                                 cudaStream_t st);
cudaError_t bep_launch_dict(const BepArgs &a, cudaStream_t st);
cudaError_t bep_launch_sweeps(const BepArgs &a, int W, cudaStream_t st);  // forward + backward; W in {1, 2, 4, 8}
// the same sweeps as one launch per height level (d_nodes: internal nodes grouped by height, lowest first)
cudaError_t bep_launch_sweeps_levels(const BepArgs &a, int W, const int32_t *d_nodes, const int32_t *level_off, int n_levels,
                                     cudaStream_t st);

// bep_reach.cu: graph assembly on the device (first pass: instance k = c + 1 forward, nPos + 1 + c backward)
cudaError_t bep_launch_reach(const uint32_t *B, const int32_t *dict, const int32_t *nsym, int mx, int n_int, int n_inst, int n_pos, int W,
                             uint32_t *A1, uint32_t *Z, uint8_t *contig, cudaStream_t st);
cudaError_t bep_launch_rows(const uint8_t *obs, const int32_t *ent_of_node, const uint32_t *Z, int n_nodes, int n_int, int n_pos,
                            uint8_t *present, cudaStream_t st);
cudaError_t bep_launch_gather(const uint32_t *B, const int32_t *rows, int n_rows, size_t row_words, uint32_t *out, cudaStream_t st);

}  // namespace bnk
