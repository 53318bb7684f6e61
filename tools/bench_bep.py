#!/usr/bin/env python3
"""Measure bnk_bep_presence on synthetic alignments with gaps and print one JSON line.

    python tools/bench_bep.py [--leaves 10000] [--cols 5000] [--gaps 0.3] [--tree balanced|caterpillar|random]

updates = edge instances x internal nodes x 2 sweeps (forward + backward).  The CPU figure beside it is the
oracle (oracle/bep.py, a literal pure-Python restatement of asr.Parsimony) on a bounded sample of instances.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--leaves", type=int, default=10000)
    ap.add_argument("--cols", type=int, default=5000)
    ap.add_argument("--gaps", type=float, default=0.3)
    ap.add_argument("--tree", default="balanced")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-budget", type=float, default=10.0)
    args = ap.parse_args()
    from bnkit_b200 import lib, synth
    if args.tree == "balanced":
        parent, dist = synth.balanced_tree(args.leaves, seed=1)
    elif args.tree == "caterpillar":
        parent, dist = synth.caterpillar_tree(args.leaves, seed=2)
    else:
        parent, dist = synth.random_tree(args.leaves, seed=5)
    n = len(parent)
    has_child = np.zeros(n, bool)
    has_child[parent[parent >= 0]] = True
    rng = np.random.default_rng(11)
    obs = np.full((n, args.cols), 255, np.uint8)
    # gaps are indel events on branches: a deleted segment is shared by every sequence below the branch, the way
    # real alignments look (independent per-sequence gaps would make almost every ancestor graph discontinuous)
    leaves = np.nonzero(~has_child)[0]
    for v in leaves:
        obs[v] = rng.integers(0, 20, size=args.cols).astype(np.uint8)
    last = np.arange(n)  # last node of each subtree (depth-first index order: a subtree is a contiguous range)
    for v in range(n - 1, 0, -1):
        last[parent[v]] = max(last[parent[v]], last[v])
    is_leaf_row = ~has_child
    target = args.gaps * len(leaves) * args.cols
    gone = 0
    while gone < target:
        v = int(rng.integers(1, n))
        ln = int(rng.geometric(1.0 / 12))
        a = int(rng.integers(0, args.cols))
        sub = np.arange(v, last[v] + 1)
        sub = sub[is_leaf_row[sub]]
        if len(sub) > len(leaves) // 4:
            continue
        blk = obs[sub, a:a + ln]
        gone += int((blk != 255).sum())
        obs[sub, a:a + ln] = 255
    tree = lib.Tree(parent, dist)
    best, info = None, None
    for _ in range(args.reps):
        t0 = time.perf_counter()
        present, inf = tree.bep_presence(obs, want_info=True)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, info = dt, inf
    n_inst = 2 * (args.cols + 1)
    n_int = int(has_child.sum())
    updates = 2.0 * n_inst * n_int
    line = {
        "metric": "edge-instance x node parsimony updates/s (BEP indel inference)", "unit": "updates/s",
        "value": updates / best, "value_device_kernels": updates / max(info[7] * 1e-6, 1e-9),
        "seconds": {"total": best, "device_sweeps_incl_copies": info[4] * 1e-6, "host_graph_assembly": info[5] * 1e-6, "device_kernels_cuda_events": info[7] * 1e-6},
        "config": {"tree": args.tree, "leaves": args.leaves, "cols": args.cols, "gap_fraction": float((obs[~has_child] == 255).mean()),
                   "instances": n_inst, "internal_nodes": n_int},
        "info": {"ancestors_without_path_after_first_pass": int(info[0]), "patch_rounds": int(info[1]),
                 "max_distinct_edge_targets": int(info[2]), "kernel_launches": int(info[3])},
        "present_fraction_ancestors": float(present[has_child].mean()),
    }
    # CPU oracle on a bounded sample of instances
    from oracle import bep
    children = bep.children_of(parent.tolist())
    is_leaf = [len(c) == 0 for c in children]
    rows = [None] * n
    for v in np.nonzero(~has_child)[0]:
        rows[v] = (obs[v] != 255).tolist()
    t0 = time.perf_counter()
    done = 0
    for i in rng.permutation(args.cols):
        inst = bep.edge_instance(rows, int(i), True, args.cols)
        bep.parsimony(children, inst, is_leaf, True)
        done += 1
        if time.perf_counter() - t0 > args.cpu_budget:
            break
    dt = time.perf_counter() - t0
    line["cpu_baseline"] = {"value": 2.0 * done * n_int / dt, "unit": "updates/s", "cores": 1, "kind": "port",
                            "sample": "%d forward instances, all %d nodes, %.1f s (pure-Python restatement of asr.Parsimony)" % (done, n, dt)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
