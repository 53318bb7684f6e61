// internal.h -- shared declarations of libbnkgpu (not part of the ABI).
#pragma once
#include "../../include/bnkgpu.h"

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <vector>

#define BNK_MAXK BNK_MAX_STATES

struct bnk_model {
    int K = 0;
    bool has_R = true;
    char name[16] = "custom";
    double F[BNK_MAXK] = {0}, prior[BNK_MAXK] = {0};
    double R[BNK_MAXK * BNK_MAXK] = {0};
    double evec[BNK_MAXK * BNK_MAXK] = {0}, ievec[BNK_MAXK * BNK_MAXK] = {0};
    double eval[BNK_MAXK] = {0};
};

namespace bnk {

int fail(int code, const char *fmt, ...);  // records the thread's last error, returns code

// bn.ctmc.GapSubstModel on the host: the base alphabet plus the gap state (K1 = K + 1 <= 21)
constexpr int GAP_MAXK = BNK_MAXK + 1;
struct GapModelHost {
    int K1 = 0;
    double mu = 0, lambda = 0;
    double F[GAP_MAXK] = {0};
    double R[GAP_MAXK * GAP_MAXK] = {0};
    double evec[GAP_MAXK * GAP_MAXK] = {0}, ievec[GAP_MAXK * GAP_MAXK] = {0};
    double eval[GAP_MAXK] = {0};
};
int indel_base_model_by_name(const char *name, bnk_model **out);
int gap_model_from_base(const bnk_model *base, double mu, double lambda, GapModelHost *g);

int model_from_rates(int K, const double *F, const double *M, bool symmetric, bool normalise, bnk_model **out);
int model_by_name(const char *name, const double *F_override, bnk_model **out);
int model_from_eigen(int K, const double *F, const double *evec, const double *ievec,
                     const double *eval, bnk_model **out);
void tree_retain(bnk_tree *t);
void tree_release(bnk_tree *t);  // deletes the tree when the last reference goes

// One step of the column-tile walk = one internal node (a node with >= 1 child in the
// static tree).  Childless nodes never get a step: their message is a table lookup made
// while folding the parent's children.  64 bytes, streamed into shared memory in chunks;
// the first two children are inline so binary trees never touch the overflow array.
struct Step {
    int32_t node;         // tree index
    int32_t parent;       // tree index or -1
    int32_t ent;          // row of this node among internal nodes (index order)
    int32_t n_child;
    int32_t slot;         // up: stack slot receiving this node's message; down: slot holding its from-above vector; -1 static root
    int32_t child_begin;  // into the ChildRef array (all children, used for index >= 2)
    int32_t c_node[2];    // first two children: tree index,
    int32_t c_slot[2];    //   up: slot of the child's message / down: slot receiving its from-above vector; -1 childless
    int32_t c_ent[2];     //   internal row or -1
    int32_t pslot;        // traceback: slot holding the parent's state (-1: none)
    int32_t oslot;        // traceback: slot receiving this node's state (-1: not needed)
    int32_t pad[2];
};
static_assert(sizeof(Step) == 64, "Step must be 64 bytes");
struct ChildRef {
    int32_t node;
    int32_t slot;
    int32_t ent;
};

}  // namespace bnk

namespace bnk {
// A walk program over a set of internal nodes (all of them, or the "narrow" top of the tree).
struct Program {
    std::vector<Step> up;          // post-order, heavy child first
    std::vector<ChildRef> up_child;
    std::vector<Step> down;        // pre-order, heavy child last
    std::vector<ChildRef> down_child;
    int32_t up_slots = 0, down_slots = 0, tb_slots = 0;
};
// One node of a wide height level (processed by the level kernels of wide.cu): binary or unary.
struct WideNode {
    int32_t node, parent, ent, n_child;
    int32_t c_node[2];  // tree index or -1
    int32_t c_ent[2];   // internal row of the child or -1 (childless)
};
static_assert(sizeof(WideNode) == 32, "WideNode must be 32 bytes");
constexpr int32_t SLOT_CHILDLESS = -1;  // Step::c_slot: the child is childless (message = table lookup)
constexpr int32_t SLOT_WIDE = -2;       // Step::c_slot: the child was handled by a wide level kernel (message in HBM)
}  // namespace bnk

struct bnk_tree {
    // reference count: bnk_tree_create = 1, every workspace holds one more, bnk_tree_destroy drops the
    // caller's -- so a tree destroyed before its workspaces stays alive until the last of them is gone
    std::atomic<int> refs{1};
    int32_t n = 0, n_int = 0, root_count = 0;
    std::vector<int32_t> parent;
    std::vector<double> dist;
    std::vector<int32_t> first, kids;  // children CSR in index order (IdxTree.children[])
    std::vector<int32_t> ent_of_node;  // -1 for childless nodes
    std::vector<int32_t> node_of_ent;
    std::vector<int32_t> height;       // 0 childless, else 1 + max over children
    bnk::Program full;                 // every internal node (general kernels; deep trees)
    // hybrid schedule: height levels 1..wide_h are wide enough to run as one launch each; the rest
    // of the tree (an upward-closed set) is walked by column-owning CTAs
    int32_t wide_h = 0;
    std::vector<std::vector<bnk::WideNode>> levels;  // levels[h - 1], h = 1..wide_h
    bnk::Program narrow;
    int32_t n_wide = 0;
    int32_t max_children = 0;
};
