// tree_plan.cpp -- host-side tree handle: validation, child lists, and the static walk
// programs the device kernels execute.
//
// The device never launches per tree level.  Alignment columns are independent
// (ThreadedDecorators runs one job per column, src/asr/ThreadedDecorators.java:28-72), so a
// CTA that owns a tile of columns can walk the whole tree by itself, keeping the live
// messages of its columns in a shared-memory stack.  The walk orders computed here bound
// that stack: the up-sweep visits the heaviest child subtree first (post-order), the
// down-sweep and the traceback visit it last (pre-order), so at most ~log2(#ancestors)
// vectors are live per column for binary trees -- independent of tree depth, which is what
// makes a 10,000-level caterpillar cost the same as a balanced tree.
//
// Input layout = dat.phylo.IdxTree (src/dat/phylo/IdxTree.java:27-32, :357-404):
// parent[i] < i (depth-first pre-order), children in index order.
#include "internal.h"

#include <algorithm>
#include <cstring>
#include <new>
#include <queue>

namespace bnk {

namespace {
struct SlotPool {  // lowest-free-index allocator
    std::priority_queue<int32_t, std::vector<int32_t>, std::greater<int32_t>> free_;
    int32_t next = 0;
    int32_t alloc()
    {
        if (!free_.empty()) { int32_t s = free_.top(); free_.pop(); return s; }
        return next++;
    }
    void release(int32_t s) { if (s >= 0) free_.push(s); }
};
}  // namespace

// Build the up / down walk programs over the internal nodes with member[v] != 0 (an upward-closed
// set: the whole tree, or its narrow top).  Internal children outside the set are marked SLOT_WIDE.
static void build_program(const bnk_tree *t, const std::vector<uint8_t> &member, Program &pg)
{
    const int32_t n = t->n;
    auto internal = [&](int32_t v) { return t->ent_of_node[v] >= 0; };
    // weight = number of member nodes in the subtree (children have larger indices)
    std::vector<int32_t> weight(n, 0);
    for (int32_t i = n - 1; i >= 0; i--) {
        if (internal(i) && member[i]) weight[i] += 1;
        if (t->parent[i] >= 0) weight[t->parent[i]] += weight[i];
    }
    // member children of each node, heaviest first (ties: index order)
    auto member_children = [&](int32_t v, std::vector<int32_t> &out) {
        out.clear();
        for (int32_t e = t->first[v]; e < t->first[v + 1]; e++)
            if (internal(t->kids[e]) && member[t->kids[e]]) out.push_back(t->kids[e]);
        std::stable_sort(out.begin(), out.end(), [&](int32_t a, int32_t b) { return weight[a] > weight[b]; });
    };
    std::vector<int32_t> up_order, down_order;
    {
        std::vector<int32_t> stack, tmp;
        std::vector<uint8_t> expanded(n, 0);
        for (int32_t r = 0; r < n; r++) {
            if (t->parent[r] >= 0 || !internal(r) || !member[r]) continue;
            stack.push_back(r);
            while (!stack.empty()) {
                const int32_t v = stack.back();
                if (expanded[v]) { stack.pop_back(); up_order.push_back(v); continue; }
                expanded[v] = 1;
                member_children(v, tmp);
                for (auto it = tmp.rbegin(); it != tmp.rend(); ++it) stack.push_back(*it);  // heaviest on top
            }
        }
    }
    {
        std::vector<int32_t> stack, tmp;
        for (int32_t r = n - 1; r >= 0; r--)
            if (t->parent[r] < 0 && internal(r) && member[r]) stack.push_back(r);
        while (!stack.empty()) {
            const int32_t v = stack.back();
            stack.pop_back();
            down_order.push_back(v);
            member_children(v, tmp);  // heaviest first -> pushed first -> popped last
            for (int32_t c : tmp) stack.push_back(c);
        }
    }
    auto child_kind = [&](int32_t u) { return !internal(u) ? SLOT_CHILDLESS : (member[u] ? 0 : SLOT_WIDE); };

    pg.up.clear(); pg.up_child.clear(); pg.down.clear(); pg.down_child.clear();
    {
        SlotPool pool;
        std::vector<int32_t> slot_of(n, -1);
        for (int32_t v : up_order) {
            Step s;
            std::memset(&s, 0, sizeof(s));
            s.node = v; s.parent = t->parent[v]; s.ent = t->ent_of_node[v];
            s.child_begin = (int32_t)pg.up_child.size();
            s.n_child = t->first[v + 1] - t->first[v];
            s.pslot = s.oslot = -1;
            for (int k = 0; k < 2; k++) { s.c_node[k] = -1; s.c_slot[k] = -1; s.c_ent[k] = -1; }
            for (int32_t e = t->first[v]; e < t->first[v + 1]; e++) {
                const int32_t u = t->kids[e];
                const int k = e - t->first[v];
                const int32_t kind = child_kind(u);
                const int32_t sl = kind == 0 ? slot_of[u] : kind;
                if (k < 2) { s.c_node[k] = u; s.c_slot[k] = sl; s.c_ent[k] = t->ent_of_node[u]; }
                pg.up_child.push_back(ChildRef{u, sl, t->ent_of_node[u]});
            }
            // the node's slot is taken BEFORE its children's are released: a step never writes the slot it reads
            // (the lane walker of sweeps_lane.cu hands messages from warp to warp with one barrier per step)
            s.slot = (t->parent[v] >= 0) ? pool.alloc() : -1;
            for (int32_t e = t->first[v]; e < t->first[v + 1]; e++) pool.release(slot_of[t->kids[e]]);
            slot_of[v] = s.slot;
            pg.up.push_back(s);
        }
        pg.up_slots = pool.next;
    }
    {
        SlotPool tpool, spool;
        std::vector<int32_t> tslot(n, -1), sslot(n, -1), remaining(n, 0);
        for (int32_t v : down_order) {
            Step s;
            std::memset(&s, 0, sizeof(s));
            s.node = v; s.parent = t->parent[v]; s.ent = t->ent_of_node[v];
            s.slot = tslot[v];
            s.child_begin = (int32_t)pg.down_child.size();
            s.n_child = t->first[v + 1] - t->first[v];
            for (int k = 0; k < 2; k++) { s.c_node[k] = -1; s.c_slot[k] = -1; s.c_ent[k] = -1; }
            int32_t n_member_child = 0;
            for (int32_t e = t->first[v]; e < t->first[v + 1]; e++) {
                const int32_t u = t->kids[e];
                const int k = e - t->first[v];
                const int32_t kind = child_kind(u);
                int32_t sl = kind;
                if (kind == 0) { sl = tpool.alloc(); tslot[u] = sl; n_member_child++; }
                if (k < 2) { s.c_node[k] = u; s.c_slot[k] = sl; s.c_ent[k] = t->ent_of_node[u]; }
                pg.down_child.push_back(ChildRef{u, sl, t->ent_of_node[u]});
            }
            tpool.release(tslot[v]);  // after the children's slots: a child never reuses its parent's slot in the same step
            // traceback slots
            const int32_t p = t->parent[v];
            s.pslot = (p >= 0) ? sslot[p] : -1;
            if (p >= 0 && --remaining[p] == 0) spool.release(sslot[p]);
            if (n_member_child > 0) { sslot[v] = spool.alloc(); remaining[v] = n_member_child; }
            s.oslot = sslot[v];
            pg.down.push_back(s);
        }
        pg.down_slots = tpool.next;
        pg.tb_slots = spool.next;
    }
}

static const int32_t WIDE_MIN_NODES = 16;  // a height level with at least this many nodes runs as one launch

static void build_plan(bnk_tree *t)
{
    const int32_t n = t->n;
    t->first.assign(n + 1, 0);
    for (int32_t i = 0; i < n; i++)
        if (t->parent[i] >= 0) t->first[t->parent[i] + 1]++;
    for (int32_t i = 0; i < n; i++) t->first[i + 1] += t->first[i];
    t->kids.assign(t->first[n], 0);
    {
        std::vector<int32_t> fill(t->first.begin(), t->first.end() - 1);
        for (int32_t i = 0; i < n; i++)
            if (t->parent[i] >= 0) t->kids[fill[t->parent[i]]++] = i;
    }
    t->ent_of_node.assign(n, -1);
    t->node_of_ent.clear();
    t->root_count = 0;
    for (int32_t i = 0; i < n; i++) {
        if (t->parent[i] < 0) t->root_count++;
        if (t->first[i + 1] > t->first[i]) {
            t->ent_of_node[i] = (int32_t)t->node_of_ent.size();
            t->node_of_ent.push_back(i);
        }
    }
    t->n_int = (int32_t)t->node_of_ent.size();
    t->max_children = 0;
    for (int32_t i = 0; i < n; i++) t->max_children = std::max(t->max_children, t->first[i + 1] - t->first[i]);

    std::vector<uint8_t> all(n, 1);
    build_program(t, all, t->full);

    // heights and the hybrid split
    t->height.assign(n, 0);
    int32_t hmax = 0;
    for (int32_t i = n - 1; i >= 0; i--) {
        if (t->ent_of_node[i] >= 0 && t->height[i] == 0) t->height[i] = 1;  // internal with childless children only
        if (t->parent[i] >= 0) t->height[t->parent[i]] = std::max(t->height[t->parent[i]], t->height[i] + 1);
        hmax = std::max(hmax, t->height[i]);
    }
    std::vector<int32_t> width(hmax + 2, 0);
    for (int32_t i = 0; i < n; i++)
        if (t->ent_of_node[i] >= 0) width[t->height[i]]++;
    // the level kernels handle nodes with at most two children: a multifurcation and everything above
    // it stays with the walker (the wide set must be closed under taking descendants)
    std::vector<uint8_t> poly(n, 0);
    for (int32_t i = n - 1; i >= 0; i--) {
        if (t->first[i + 1] - t->first[i] > 2) poly[i] = 1;
        if (poly[i] && t->parent[i] >= 0) poly[t->parent[i]] = 1;
    }
    std::fill(width.begin(), width.end(), 0);
    for (int32_t i = 0; i < n; i++)
        if (t->ent_of_node[i] >= 0 && !poly[i] && t->parent[i] >= 0) width[t->height[i]]++;
    t->wide_h = 0;
    while (t->wide_h + 1 <= hmax && width[t->wide_h + 1] >= WIDE_MIN_NODES) t->wide_h++;
    // the root of a tree is never wide (its level has one node), so the narrow set is never empty
    std::vector<uint8_t> narrow(n, 0);
    t->levels.assign(t->wide_h, {});
    t->n_wide = 0;
    for (int32_t i = 0; i < n; i++) {
        if (t->ent_of_node[i] < 0) continue;
        if (t->height[i] > t->wide_h || t->parent[i] < 0 || poly[i]) { narrow[i] = 1; continue; }
        WideNode w;
        w.node = i; w.parent = t->parent[i]; w.ent = t->ent_of_node[i];
        w.n_child = t->first[i + 1] - t->first[i];
        for (int k = 0; k < 2; k++) {
            const bool has = k < w.n_child;
            const int32_t u = has ? t->kids[t->first[i] + k] : -1;
            w.c_node[k] = u;
            w.c_ent[k] = has ? t->ent_of_node[u] : -1;
        }
        t->levels[t->height[i] - 1].push_back(w);
        t->n_wide++;
    }
    // narrow must be upward closed: a wide node's parent has a larger height, hence is narrow or wide
    // at a higher level -- fine; but a root that was forced narrow may sit below wide_h: its
    // descendants are still wide (downward closed), which is all the kernels need.
    build_program(t, narrow, t->narrow);
}

}  // namespace bnk

using namespace bnk;

extern "C" int bnk_tree_create(int32_t n_nodes, const int32_t *parent, const double *dist, bnk_tree **out)
{
    if (!out || !parent) return fail(BNK_ERR_ARG, "tree: null argument");
    if (n_nodes < 1) return fail(BNK_ERR_TREE, "tree: needs at least one node");
    for (int32_t i = 0; i < n_nodes; i++) {
        if (parent[i] >= i || parent[i] < -1)
            return fail(BNK_ERR_TREE, "tree: parent[%d] = %d violates parent < child (IdxTree depth-first order)", i, parent[i]);
        if (dist && parent[i] >= 0 && !(dist[i] >= 0.0))
            return fail(BNK_ERR_TREE, "tree: distance[%d] is negative or NaN", i);
    }
    bnk_tree *t = new (std::nothrow) bnk_tree();
    if (!t) return fail(BNK_ERR_NOMEM, "tree: out of host memory");
    t->n = n_nodes;
    t->parent.assign(parent, parent + n_nodes);
    if (dist) t->dist.assign(dist, dist + n_nodes);
    else t->dist.assign(n_nodes, 1.0);
    build_plan(t);
    *out = t;
    return BNK_OK;
}

namespace bnk {
void tree_retain(bnk_tree *t) { if (t) t->refs.fetch_add(1, std::memory_order_relaxed); }
void tree_release(bnk_tree *t)
{
    if (t && t->refs.fetch_sub(1, std::memory_order_acq_rel) == 1) delete t;
}
}  // namespace bnk
extern "C" void bnk_tree_destroy(bnk_tree *t) { tree_release(t); }
extern "C" int bnk_tree_size(const bnk_tree *t) { return t ? t->n : fail(BNK_ERR_ARG, "tree: null"); }
extern "C" int bnk_tree_n_internal(const bnk_tree *t) { return t ? t->n_int : fail(BNK_ERR_ARG, "tree: null"); }

// IdxTree.getPrunedIndex(pruneMe, true) (IdxTree.java:411-464) on [node][col] masks.
// A node survives iff it is present and reachable, through present nodes only, from a
// present leaf by walking up and then flooding down from the top of that walk.
extern "C" int bnk_tree_prune_orphans(const bnk_tree *t, int32_t n_cols, const uint8_t *in, uint8_t *out)
{
    if (!t || !in || !out || n_cols < 0) return fail(BNK_ERR_ARG, "prune: bad argument");
    const int32_t n = t->n;
    std::vector<uint8_t> keep(n), pres(n);
    std::vector<int32_t> stack;
    for (int32_t c = 0; c < n_cols; c++) {
        for (int32_t v = 0; v < n; v++) { pres[v] = in[(size_t)v * n_cols + c] != 0; keep[v] = 0; }
        for (int32_t leaf = 0; leaf < n; leaf++) {
            if (t->first[leaf + 1] > t->first[leaf]) continue;
            int32_t top = -1, idx = leaf;
            while (pres[idx] && !keep[idx]) {
                keep[idx] = 1;
                top = idx;
                idx = t->parent[idx];
                if (idx < 0) break;
            }
            if (top < 0) continue;
            stack.assign(1, top);
            while (!stack.empty()) {
                const int32_t v = stack.back();
                stack.pop_back();
                if (!pres[v]) continue;
                for (int32_t e = t->first[v]; e < t->first[v + 1]; e++) {
                    const int32_t u = t->kids[e];
                    if (keep[u] || !pres[u]) continue;
                    keep[u] = 1;
                    stack.push_back(u);
                }
            }
        }
        for (int32_t v = 0; v < n; v++) out[(size_t)v * n_cols + c] = (uint8_t)(pres[v] && keep[v]);
    }
    return BNK_OK;
}
