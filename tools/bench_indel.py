"""Gap-augmented peeling at C3 size: one IndelPeeler.optimiseMuLambda evaluation = new
GapSubstModel + per-branch tables + peel of all columns.  Prints one JSON line; used for the ncu captures in profiles/.
  python tools/bench_indel.py [--tree balanced|caterpillar|random] [--leaves N] [--cols C] [--evals E] [--parity-cols P]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tree", default="balanced")
    ap.add_argument("--leaves", type=int, default=10000)
    ap.add_argument("--cols", type=int, default=5000)
    ap.add_argument("--evals", type=int, default=5)
    ap.add_argument("--gaps", type=float, default=0.3)
    ap.add_argument("--parity-cols", type=int, default=8)
    ap.add_argument("--brent", action="store_true", help="also run the whole optimiseMuLambda search")
    args = ap.parse_args()
    from bnkit_b200 import lib, synth
    if args.tree == "balanced":
        parent, dist = synth.balanced_tree(args.leaves, seed=1)
    elif args.tree == "caterpillar":
        parent, dist = synth.caterpillar_tree(args.leaves, seed=2)
    else:
        parent, dist = synth.random_tree(args.leaves, seed=5)
    n = len(parent)
    rng = np.random.default_rng(4)
    leaf = np.ones(n, bool)
    leaf[parent[parent >= 0]] = False
    base = rng.integers(0, 20, (int(leaf.sum()), 1)).astype(np.uint8)
    noise = rng.integers(0, 20, (int(leaf.sum()), args.cols)).astype(np.uint8)
    obs = np.full((n, args.cols), 255, np.uint8)
    obs[leaf] = np.where(rng.random(noise.shape) < 0.2, noise, base)
    size = np.ones(n, np.int64)
    for v in range(n - 1, 0, -1):
        size[parent[v]] += size[v]
    for c in range(args.cols):   # clade-wise gaps
        g = 0
        while g < args.gaps * leaf.sum():
            v = int(rng.integers(1, n))
            if size[v] > n // 8:
                continue
            obs[v:v + size[v], c] = 255
            g += (size[v] + 1) // 2
    m = lib.Model("LG")
    tree = lib.Tree(parent, dist)
    h = lib.Indel(m, tree, args.cols)
    h.set_alignment(obs)
    geom = 1.0 / 3000
    ms, pt, pp = [], [], []
    for i in range(args.evals + 2):
        t0 = time.perf_counter()
        h.set_rates(0.02 + 0.001 * i, 0.02 + 0.001 * i)
        _, s = h.column_logl(1.0, geom)
        if i >= 2:
            ms.append((time.perf_counter() - t0) * 1e3)
            pt.append(h.phase_ms(0))
            pp.append(h.phase_ms(1))
    ni = tree.n_internal
    out = {"workload": "%s %d leaves x %d columns, LG + gap state, %.0f %% clade-wise gaps" % (args.tree, args.leaves, args.cols, 100 * args.gaps),
           "ms_per_evaluation_wall": float(np.mean(ms)), "tables_ms": float(np.mean(pt)), "peel_ms": float(np.mean(pp)),
           "updates_per_s": ni * args.cols / (np.mean(pp) * 1e-3),
           "algorithmic_GB": 2.0 * (21 * 8 + 9) * ni * (args.cols + 1) / 1e9,
           "launches_per_evaluation": int(h.launch_count() // (args.evals + 2)), "sum_logl": s}
    out["achieved_GBps"] = out["algorithmic_GB"] / (out["peel_ms"] * 1e-3)
    if args.parity_cols > 0:
        from oracle import oracle as O
        O.build()
        sel = np.linspace(0, args.cols - 1, args.parity_cols).astype(int)
        h.set_rates(0.02, 0.02)
        got, _ = h.column_logl(1.0, geom)
        t0 = time.perf_counter()
        want = O.indel_column_logl(O.GapModel(O.Model("LG"), 0.02, 0.02), parent, dist, np.ascontiguousarray(obs[:, sel]), 1.0, geom, nthreads=args.parity_cols)
        dt = time.perf_counter() - t0
        out["parity"] = {"cols": int(len(sel)), "logl_max_rel": float(np.max(np.abs(got[sel] - want) / np.abs(want)))}
        out["cpu_oracle_updates_per_s"] = ni * len(sel) / dt
        out["cpu_oracle_threads"] = int(args.parity_cols)
    if args.brent:
        t0 = time.perf_counter()
        x, ne = h.optimise_mulambda(1e-4, 0.5, geom)
        out["brent"] = {"mu_lambda": x, "likelihood_evaluations": ne, "seconds": time.perf_counter() - t0}
    h.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
