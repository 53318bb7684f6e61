"""Probe of the host-buffer marginal path: wall time of bnk_marginal (all ancestors, pinned buffers) on C3a for several
column-chunk sizes of the pipeline, against the box's pinned D2H rate for a contiguous and for a pitched copy of the
same geometry.  python tools/probe_e2e.py"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from bnkit_b200 import lib, synth

parent, dist = synth.balanced_tree(10000, seed=1)
n = len(parent); n_cols = 5000
rng = np.random.default_rng(4)
r4 = synth.gamma_category_rates(0.5, 4)
rates = r4[rng.integers(0, 4, n_cols)]
leaf = np.ones(n, bool); leaf[parent[parent >= 0]] = False
obs = np.full((n, n_cols), 255, np.uint8)
obs[leaf] = rng.integers(0, 20, (int(leaf.sum()), n_cols))
m = lib.Model("LG"); tree = lib.Tree(parent, dist); n_int = tree.n_internal
h_obs = torch.from_numpy(obs).pin_memory()
h_post = torch.empty((n_int, n_cols, 20), dtype=torch.float64).pin_memory()
h_logl = torch.empty((n_cols,), dtype=torch.float64).pin_memory()
out = {}
L = lib.lib()
for chunk in (0, 2500, 1250, 625, 313):
    ws = lib.Workspace(m, tree, n_cols)
    ws.set_rates(n_cols, rates)
    if chunk: ws.set_pipeline(chunk)
    ts = []
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        lib.check(L.bnk_marginal(ws.h, n_cols, h_obs.data_ptr(), None, None, -1, h_post.data_ptr(), h_logl.data_ptr()))
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    out["marginal_ms_chunk_%d" % chunk] = min(ts[1:])
    ws.close()
d = torch.empty((n_int, 1250, 20), dtype=torch.float64, device="cuda")
for name, dst in (("contiguous", h_post.view(-1)[: d.numel()].view(n_int, 1250, 20)), ("pitched", h_post[:, :1250, :])):
    ts = []
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dst.copy_(d, non_blocking=True)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    out["d2h_GBps_" + name] = d.numel() * 8 / 1e9 / min(ts[1:])
print(json.dumps(out))
