#!/usr/bin/env python3
"""End-to-end ASR time on BASELINE.json configs[0] and [1] (the reference's own data files, as committed in
tests/golden/): indel inference (BEP) -> orphan pruning -> joint reconstruction (C1: data/66.*, LG) or marginal
reconstruction of every ancestor (C2: data/3_2_1_1_filt.*, JTT), through the C ABI with host buffers.  The CPU
column beside it is the oracle chain (pure-Python BEP + the threaded C restatement of VarElim) on the same inputs.
Prints one JSON line per config."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
AA = "ARNDCQEGHILKMFPSTWYV"


def load(name):
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", name)))
    parent = np.asarray(fx["tree"]["Parents"], np.int32)
    dist = np.asarray(fx["tree"]["Distances"], np.float64)
    labels = fx["tree"]["Labels"]
    n, C = len(parent), len(fx["seqs"][0])
    obs = np.full((n, C), 255, np.uint8)
    for nm, seq in zip(fx["names"], fx["seqs"]):
        v = labels.index(nm)
        for c, ch in enumerate(seq):
            if ch != "-":
                obs[v, c] = AA.index(ch)
    return parent, dist, obs


def main():
    from bnkit_b200 import lib
    from oracle import bep, oracle as O
    for tag, fixture, model, mode in (("C1 data/66 joint LG", "c1_66.json", "LG", "joint"),
                                      ("C2 data/3_2_1_1_filt marginal (all ancestors) JTT", "c2_3_2_1_1_filt.json", "JTT", "marginal")):
        parent, dist, obs = load(fixture)
        n, C = obs.shape
        tree = lib.Tree(parent, dist)
        m = lib.Model(model)
        ws = lib.Workspace(m, tree, C, device=0)

        def gpu_once():
            t0 = time.perf_counter()
            present = tree.prune_orphans(tree.bep_presence(obs))
            t1 = time.perf_counter()
            ws.set_rates(C, None)
            res = ws.joint(obs, present) if mode == "joint" else ws.marginal(obs, present)[0]
            t2 = time.perf_counter()
            return present, res, t1 - t0, t2 - t1
        gpu_once()  # warm-up (allocations, first launches)
        best = None
        for _ in range(5):
            present, res, tb, ti = gpu_once()
            if best is None or tb + ti < best[0] + best[1]:
                best = (tb, ti)
        # CPU oracle chain
        has_child = np.zeros(n, bool)
        has_child[parent[parent >= 0]] = True
        rows = [None if has_child[v] else (obs[v] != 255).tolist() for v in range(n)]
        t0 = time.perf_counter()
        mask, _ = bep.presence_mask(parent.tolist(), rows, C)
        mask = np.asarray(mask, np.uint8)
        pm = np.stack([O.prune(parent, np.ascontiguousarray(mask[:, c]), True)[0] for c in range(C)], axis=1)
        t1 = time.perf_counter()
        om = O.Model(model)
        threads = os.cpu_count() or 1
        if mode == "joint":
            want = O.fast_joint(om, parent, dist, obs, pm, None, nthreads=threads)
            same = bool(np.array_equal(res, want))
        else:
            want, _ = O.fast_marginal(om, parent, dist, obs, pm, None, nthreads=threads)
            wi = want[has_child]
            same = bool(np.array_equal(np.isnan(res), np.isnan(wi)) and np.allclose(res, wi, rtol=1e-9, atol=1e-12, equal_nan=True))
        t2 = time.perf_counter()
        n_int = int(has_child.sum())
        print(json.dumps({"config": tag, "leaves": int((~has_child).sum()), "ancestors": n_int, "columns": C,
                          "gap_fraction": float((obs[~has_child] == 255).mean()),
                          "b200_ms": {"indel_inference_bep": best[0] * 1e3, "character_inference": best[1] * 1e3,
                                      "total": (best[0] + best[1]) * 1e3},
                          "cpu_oracle_ms": {"indel_inference_bep_python": (t1 - t0) * 1e3,
                                            "character_inference_c_%d_threads" % threads: (t2 - t1) * 1e3},
                          "results_equal_to_oracle": same and bool(np.array_equal(present, pm))}))
        ws.close()


if __name__ == "__main__":
    main()
