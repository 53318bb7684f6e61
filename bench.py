#!/usr/bin/env python3
"""bench.py -- throughput of the GRASP per-column tree-inference path on B200.

One step = one pass of the hot path over the batch of synthetic columns:
    P(t)/log P(t) table build  +  joint reconstruction (max-product up-sweep + traceback)
    +  marginal reconstruction of ALL ancestors (sum-product up-sweep + down-sweep) + log-likelihoods
on BASELINE.json configs[2] (C3a): 10,000-leaf balanced tree x 5,000 columns, LG + Gamma4 (per-column category
rates), FP64.  Metric = column x node partial updates / s, one update = one (column, internal node) message
computation in one sweep: 3 sweeps per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3a|c3b|c4|small] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU).  The FIXED problem is sharded: rank r owns a contiguous block
of 5000 / N columns ("scaling": "strong"); no sweep contains a data-path collective.  The same JSON line
carries sub-results measured the same way (barrier + CUDA events, max over ranks) for the other BASELINE
configurations: C3b (caterpillar), C4 (100,000 leaves x 2,000 columns, WAG, sharded) and the C5 optimisation
loop -- per evaluation a rate change, a table rebuild, a log-likelihood sweep and the NCCL all-reduce of the
per-rank sums INSIDE the timed region, plus one gather of reconstructed states.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes per update, SURVEY.md section 8(d) (store-all-partials schedule, FP64, K = 20)
BYTES_PER_UPDATE = {"joint_up+traceback": 361.0, "sum_up": 337.0, "sum_down": 832.0}
# Bytes the r2 kernels no longer move for a height-1 node under a level-scheduled parent ("cherry-direct" + the fused
# down-sweep), per (node, column): its message row is neither written by an expansion kernel nor
# read by the parent (2 x 160 B; joint: + 2 x 20 B of pointer rows; the 2-byte record index is written once and read
# by the parent / the traceback instead), its from-above vector is neither written nor read (2 x 160 B) and the
# parent's down-sweep does not read its message (160 B).  The algorithmic bytes of the sweeps are restated by these.
BYTES_SAVED_PER_DIRECT = {"joint_up+traceback": 336.0, "sum_up": 316.0, "sum_down": 480.0}
TABLE_BYTES = 3200.0  # per (branch, category) per sweep

WORKLOADS = {
    "c3a": dict(tree="balanced", leaves=10000, cols=5000, model="LG", ncat=4, name="C3a balanced 10k-leaf x 5k-col LG+G4"),
    "c3b": dict(tree="caterpillar", leaves=10000, cols=5000, model="LG", ncat=4, name="C3b caterpillar 10k-leaf x 5k-col LG+G4"),
    "small": dict(tree="balanced", leaves=512, cols=1024, model="LG", ncat=4, name="small 512-leaf x 1024-col LG+G4"),
    # BASELINE configs[3]: 100k-leaf random tree x 2k columns, WAG, one rate
    "c4": dict(tree="random", leaves=100000, cols=2000, model="WAG", ncat=1, name="C4 random 100k-leaf x 2k-col WAG"),
}
CPU_SAMPLE_COLS = 256  # columns of the CPU arm's sample (64 per rate category): the same for cpu_baseline and --impl reference


def model_parts(name, impl):
    """eigensystem / prior of the model for the alignment simulator: from the library in the B200 arm, from the
    oracle in the reference arm (which must not map libbnkgpu.so at all)"""
    if impl == "reference":
        from oracle import oracle as O
        return O.Model(name).parts()
    from bnkit_b200 import lib
    return lib.Model(name).parts()


def make_workload(spec, rank=0, cols=None, impl="b200"):
    """the FIXED problem of a workload (same on every rank): (model name, parent, dist, rates, obs [n][cols])"""
    from bnkit_b200 import synth  # workload generator only (numpy)
    leaves = spec["leaves"]
    n_cols = cols or spec["cols"]
    if spec["tree"] == "balanced":
        parent, dist = synth.balanced_tree(leaves, seed=1)
    elif spec["tree"] == "caterpillar":
        parent, dist = synth.caterpillar_tree(leaves, seed=2)
    else:
        parent, dist = synth.random_tree(leaves, seed=5)
    parts = model_parts(spec["model"], impl)
    K = len(parts["prior"])
    cat_rates = synth.gamma_category_rates(0.5, spec["ncat"]) if spec["ncat"] > 1 else np.ones(1)
    rng = np.random.default_rng(4)
    rates = cat_rates[rng.integers(0, spec["ncat"], n_cols)]
    if leaves <= 20000:
        obs = synth.simulate_alignment(parent, dist, parts, n_cols, rates, seed=3)
    else:  # very large trees: a conserved base state per leaf with 20 % noise (numpy simulation would take minutes)
        leaf = np.ones(len(parent), dtype=bool)
        leaf[parent[parent >= 0]] = False
        nl = int(leaf.sum())
        base = rng.integers(0, K, (nl, 1)).astype(np.uint8)
        noise = rng.integers(0, K, (nl, n_cols)).astype(np.uint8)
        obs = np.full((len(parent), n_cols), 255, dtype=np.uint8)
        obs[leaf] = np.where(rng.random((nl, n_cols)) < 0.2, noise, base)
    del rank
    return spec["model"], parent, dist, rates, obs


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def n_internal(parent):
    return int((np.bincount(parent[parent >= 0], minlength=len(parent)) > 0).sum())


def sample_columns(rates, per_cat, seed=0):
    rng = np.random.default_rng(seed)
    sel = []
    for r in np.unique(rates):
        idx = np.nonzero(rates == r)[0]
        sel.extend(rng.choice(idx, size=min(per_cat, len(idx)), replace=False).tolist())
    return np.array(sorted(sel), dtype=np.int64)


def cpu_oracle_rate(model_name, parent, dist, rates, obs, threads, n_sample=CPU_SAMPLE_COLS):
    """Times the CPU restatement (oracle/) on a fixed column sample; returns (updates/s, sample text, seconds)."""
    from oracle import oracle as O
    O.build()
    om = O.Model(model_name)
    sel = sample_columns(rates, max(1, n_sample // max(1, len(np.unique(rates)))))
    o = np.ascontiguousarray(obs[:, sel])
    r = np.ascontiguousarray(rates[sel])
    t0 = time.perf_counter()
    O.fast_joint(om, parent, dist, o, None, r, nthreads=threads)
    O.fast_marginal(om, parent, dist, o, None, r, nthreads=threads, want_post=True)
    t = time.perf_counter() - t0
    nc = len(sel)
    return 3.0 * n_internal(parent) * nc / t, "%d of %d columns (%d per rate category), all %d nodes, joint + marginal(all ancestors), %d threads, %.1f s" % (
        nc, obs.shape[1], nc // max(1, len(np.unique(rates))), len(parent), threads, t), t


def java_probe():
    """BASELINE.md section 3 asks for bnkit's own JVM path when a JDK exists on the box: probe and say so."""
    exe = shutil.which("java")
    if not exe:
        return "absent (no `java` on PATH): the CPU arm is the C restatement of the reference's algorithm (kind: port)"
    try:
        out = subprocess.run([exe, "-version"], capture_output=True, text=True, timeout=20)
        return "present (%s) but /root/reference is not on the GPU box and no bnkit.jar ships: CPU arm stays the port" % (
            (out.stderr or out.stdout).strip().splitlines()[0])
    except Exception as e:  # pragma: no cover
        return "probe failed: %r" % (e,)


def host_memory_available():
    """Bytes of host memory this process may still take: free memory, capped by the cgroup limit if there is one."""
    avail = None
    try:
        import psutil
        avail = int(psutil.virtual_memory().available)
    except Exception:
        pass
    try:
        mx = open("/sys/fs/cgroup/memory.max").read().strip()
        if mx != "max":
            cur = int(open("/sys/fs/cgroup/memory.current").read().strip())
            lim = int(mx) - cur
            avail = lim if avail is None else min(avail, lim)
    except Exception:
        pass
    return avail if avail is not None else 1 << 62


class Shard:
    """One rank's device-resident share of a workload: workspace + buffers for its column block."""

    def __init__(self, lib, torch, spec, model_name, parent, dist, rates, obs, world, rank, local_rank, weak_cols=0,
                 want_post=True, present=None):
        from bnkit_b200 import dist as bd
        self.lib, self.torch = lib, torch
        self.dev = torch.device("cuda", local_rank)
        self.model = lib.Model(model_name)
        self.tree = lib.Tree(parent, dist)
        self.n, self.n_int, self.K = self.tree.n, self.tree.n_internal, self.model.K
        self.total_cols = obs.shape[1]
        if weak_cols:  # the old weak-scaling layout: every rank runs the whole problem
            self.lo, self.hi = 0, self.total_cols
        else:
            self.lo, self.hi = bd.shard_columns(self.total_cols, world, rank)
        self.ncols = self.hi - self.lo
        self.dist0 = np.ascontiguousarray(dist, dtype=np.float64)
        self.rates = np.ascontiguousarray(rates[self.lo:self.hi])
        self.obs_local = np.ascontiguousarray(obs[:, self.lo:self.hi])
        self.present_local = None if present is None else np.ascontiguousarray(present[:, self.lo:self.hi])
        nc = max(self.ncols, 1)
        self.ws = lib.Workspace(self.model, self.tree, nc, device=local_rank, stream=torch.cuda.current_stream().cuda_stream)
        # GRASP instantiates extants only (POGTree.getNodeInstance): the resident loop asserts it instead of
        # paying a scan + host round trip per call; the host-buffer (e2e) calls below run with AUTO
        self.ws.set_evidence_mode(lib.EVIDENCE_LEAVES)
        self.ws.set_rates(nc, self.rates if self.ncols else None)
        self.d_obs = torch.from_numpy(self.obs_local if self.ncols else np.full((self.n, 1), 255, np.uint8)).to(self.dev)
        self.d_present = None if self.present_local is None else torch.from_numpy(self.present_local).to(self.dev)
        self.d_state = torch.empty((self.n, nc), dtype=torch.uint8, device=self.dev)
        self.d_post = torch.empty((self.n_int if want_post else 1, nc, self.K), dtype=torch.float64, device=self.dev)
        self.d_logl = torch.empty((nc,), dtype=torch.float64, device=self.dev)
        self.d_sum = torch.zeros((1,), dtype=torch.float64, device=self.dev)
        self.serial = bool(os.environ.get("BNK_BENCH_SERIAL"))

    def step(self):
        """tables (marked stale by the distance upload, as in an optimisation step) + joint + marginal, as ONE call:
        bnk_joint_marginal_dev runs the two independent passes side by side on two streams"""
        if self.ncols == 0:
            return
        if self.serial:
            return self.step_serial()
        self.ws.set_distances(self.dist0)
        self.ws.joint_marginal_dev(self.ncols, self.d_obs, self.d_present, self.d_state, self.d_post, self.d_logl)

    def step_serial(self):
        """the same work as two calls, one pass after the other (per-phase CUDA-event times are only meaningful here)"""
        if self.ncols == 0:
            return
        self.ws.set_distances(self.dist0)
        self.ws.joint_dev(self.ncols, self.d_obs, self.d_present, self.d_state)
        self.ws.marginal_dev(self.ncols, self.d_obs, self.d_present, self.d_post, self.d_logl)

    def close(self):
        self.ws.close()
        for k in ("d_obs", "d_present", "d_state", "d_post", "d_logl", "d_sum"):
            setattr(self, k, None)
        self.torch.cuda.empty_cache()


def fused_height1_nodes(parent):
    """height-1 nodes whose parent is a level-scheduled node (height <= wide_h, not a root): the ones down_wide2_kernel
    evaluates in the parent's thread (same rule as tree_plan.cpp / api.cu on binary trees: levels with >= 16 nodes)"""
    parent = np.asarray(parent)
    n = len(parent)
    height = np.zeros(n, dtype=np.int64)
    nch = np.zeros(n, dtype=np.int64)
    for v in range(n - 1, 0, -1):   # children have larger indices than their parents (IdxTree order)
        p = parent[v]
        if p >= 0:
            height[p] = max(height[p], height[v] + 1)
            nch[p] += 1
    if nch.max() > 2:
        return 0
    internal = nch > 0
    width = np.bincount(height[internal & (parent >= 0)], minlength=3)
    wide_h = 0
    while wide_h + 1 < len(width) and width[wide_h + 1] >= 16:
        wide_h += 1
    if wide_h < 2:
        return 0
    par = parent.copy()
    par[par < 0] = 0
    ok = (height == 1) & (parent >= 0) & (height[par] >= 2) & (height[par] <= wide_h) & (parent[par] >= 0)
    return int(ok.sum())


def timed_steps(torch, dist_, world, fn, warmup, steps, after_step=None):
    """W untimed steps, then exactly K steps between barrier + synchronize, CUDA events on the launching
    stream; returns this rank's milliseconds (the caller takes the max over ranks)"""
    def barrier():
        if world > 1:
            dist_.barrier()
        torch.cuda.synchronize()
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
        if after_step:
            after_step()
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def max_over_ranks(torch, dist_, world, dev, values):
    if world == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=dev)
    dist_.all_reduce(t, op=dist_.ReduceOp.MAX)
    return [float(x) for x in t]


def parity_sample(torch, shard, model_name, parent, dist, per_cat, query=None):
    """compare a column sample of the rank's own device buffers (as the timed loop left them) with the oracle"""
    from oracle import oracle as O
    O.build()
    om = O.Model(model_name)
    sel = sample_columns(shard.rates, per_cat)
    tsel = torch.from_numpy(sel).to(shard.dev)
    st = shard.d_state[:, tsel].cpu().numpy()
    post = shard.d_post[:, tsel, :].cpu().numpy()
    logl = shard.d_logl[tsel].cpu().numpy()
    o = np.ascontiguousarray(shard.obs_local[:, sel])
    p = None if shard.present_local is None else np.ascontiguousarray(shard.present_local[:, sel])
    r = np.ascontiguousarray(shard.rates[sel])
    threads = os.cpu_count() or 1
    want = O.fast_joint(om, parent, dist, o, p, r, nthreads=threads)
    wpost, wlogl = O.fast_marginal(om, parent, dist, o, p, r, nthreads=threads)
    internal = shard.tree.internal_nodes()
    wp = wpost[internal if query is None else query]
    nan_same = bool(np.array_equal(np.isnan(post), np.isnan(wp)))
    ok = ~np.isnan(wp) & (np.abs(wp) >= 1e-300) & ~np.isnan(post)
    rel = float((np.abs(post[ok] - wp[ok]) / np.abs(wp[ok])).max()) if ok.any() else 0.0
    lrel = float((np.abs(logl - wlogl) / np.abs(wlogl)).max())
    return {"cols": int(len(sel)), "nodes": int(len(parent)), "joint_equal": bool(np.array_equal(st, want)),
            "joint_mismatches": int((st != want).sum()), "post_nan_pattern_equal": nan_same,
            "post_max_rel": rel, "logl_max_rel": lrel, "against": "oracle/ (C restatement of VarElim), same inputs"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3a", choices=sorted(WORKLOADS))
    ap.add_argument("--cols", type=int, default=0, help="columns of the problem (default: the workload's)")
    ap.add_argument("--leaves", type=int, default=0, help="leaves of the tree (default: the workload's; smaller trees for profiling)")
    ap.add_argument("--weak", action="store_true", help="every rank runs the whole problem (r1's weak-scaling layout)")
    ap.add_argument("--tile-cols", type=int, default=0)
    ap.add_argument("--cpt", type=int, default=0)
    ap.add_argument("--gaps", type=float, default=0.0, help="gap fraction at the leaves; > 0 adds the BEP presence mask (general kernels)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the C3b / C4 / C5 sub-results")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--serial", action="store_true", help="time the step as two calls (bnk_joint_dev, bnk_marginal_dev) instead of bnk_joint_marginal_dev")
    ap.add_argument("--sub", default="c3b,c4,c5,indel", help="which sub-results to run")
    args = ap.parse_args()

    # stdout carries exactly one JSON line: libraries that write to file descriptor 1 on their own (NCCL prints
    # its version banner there) are sent to stderr until the line is ready
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    spec = dict(WORKLOADS[args.workload])
    if args.leaves:
        spec["leaves"] = args.leaves
        spec["name"] += " (reduced to %d leaves)" % args.leaves
    n_cols = args.cols or spec["cols"]
    threads = os.cpu_count() or 1
    config = {"workload": spec["name"], "tree": spec["tree"], "leaves": spec["leaves"], "cols": n_cols,
              "model": spec["model"] + "+G%d" % spec["ncat"], "passes": "tables + joint(up+traceback) + marginal(up+down, all ancestors)",
              "sharding": ("every rank runs all %d columns" % n_cols) if args.weak else
                          "the fixed %d columns in contiguous blocks of %d / N per GPU; no data-path collective in a sweep" % (n_cols, n_cols),
              "schedule": "wide height levels as level launches (one column per thread, messages through HBM) + column-owned walker for the narrow top",
              "l2": "per-step working set (messages 8 GB + posteriors 8 GB + pointers 2 GB at C3) >> 126 MB L2; no flush needed"}
    scaling = "weak" if args.weak else "strong"

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        model_name, parent, dist, rates, obs = make_workload(spec, 0, n_cols, impl="reference")
        vals, sample = [], ""
        for i in range(args.warmup + args.steps):
            v, sample, _ = cpu_oracle_rate(model_name, parent, dist, rates, obs, threads)
            if i >= args.warmup:
                vals.append(v)
        value = float(np.mean(vals))
        line = {"impl": "reference", "metric": "column x node partial updates/s", "value": value, "unit": "updates/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * 3.0 * n_internal(parent) * n_cols / value, "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": "updates/s", "cores": threads, "kind": "port", "sample": sample,
                                 "note": "C restatement of bnkit's VarElim path (oracle/); every step times the SAME sample "
                                         "as the B200 arm's cpu_baseline; ms_per_step = the whole workload at that rate",
                                 "jvm": java_probe()},
                "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist_
    from bnkit_b200 import lib
    from bnkit_b200 import dist as bd
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libbnkgpu has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    # one explicit stream for everything that is timed: the workspaces launch on it (its handle is passed to
    # bnk_ws_create), torch's events are recorded on it, NCCL calls are ordered after it.  (torch's default stream
    # has the handle 0 = "no stream given"; r1 / r2_v1 lines were timed that way: kernels on a private stream,
    # events on the default one, the step's own synchronising upload keeping the two within 2 % of each other.)
    torch.cuda.set_stream(torch.cuda.Stream(dev))

    model_name, parent, dist, rates, obs = make_workload(spec, rank, n_cols)
    present = None
    if args.gaps > 0:  # per-column forests: gaps at the leaves, presence mask from the device BEP + orphan pruning
        g_rng = np.random.default_rng(77)
        leaf = np.ones(len(parent), dtype=bool)
        leaf[parent[parent >= 0]] = False
        obs = obs.copy()
        obs[(g_rng.random(obs.shape) < args.gaps) & leaf[:, None]] = 255
        t_ = lib.Tree(parent, dist)
        present = t_.prune_orphans(t_.bep_presence(obs, device=local_rank))
        config["gaps"] = args.gaps
        config["mask"] = "bnk_bep_presence (bi-directional edge parsimony) + bnk_tree_prune_orphans"
        args.no_cpu = True
    sh = Shard(lib, torch, spec, model_name, parent, dist, rates, obs, world, rank, local_rank, weak_cols=args.weak,
               present=present)
    if args.tile_cols or args.cpt:
        sh.ws.configure(args.tile_cols, args.cpt)
    n, n_int, K = sh.n, sh.n_int, sh.K
    if args.serial:
        sh.serial = True
        os.environ["BNK_BENCH_SERIAL"] = "1"

    phase_ms = np.zeros(5)

    def collect_phases():
        sh.ws.sync()
        phase_ms[:] += [sh.ws.phase_ms(i) for i in range(5)]

    for _ in range(args.warmup):
        sh.step()
    if world > 1:
        dist_.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    serial_ms = None
    if not sh.serial:  # first the sweeps one after the other, for the per-phase times the roofline is computed from;
        # the timed loop follows, so the parity sample below reads what the TIMED calls left in the buffers
        serial_ms = timed_steps(torch, dist_, world, sh.step_serial, 1, args.steps, after_step=collect_phases)
        sh.step()
    launches0 = sh.ws.launch_count()
    total_ms = timed_steps(torch, dist_, world, sh.step, 0, args.steps, after_step=collect_phases if sh.serial else None)
    launches = sh.ws.launch_count() - launches0
    if serial_ms is None:
        serial_ms = total_ms
    clocks = sampler.stop() if rank == 0 else None
    phase_ms /= args.steps

    # ---- parity of what the timed loop left in the device buffers (rank 0's block)
    parity = None
    if rank == 0 and not args.no_parity and sh.ncols > 0:
        parity = parity_sample(torch, sh, model_name, parent, dist, 16)

    # ---- end to end through the host-buffer ABI (pinned host memory, copies inside the timed region)
    e2e = None
    if not args.no_e2e and sh.ncols > 0:
        nc = sh.ncols
        # the leg pins n_int x ncols x K doubles of host memory per rank (8 GB / N at C3); if the box (or its cgroup)
        # cannot hold that, the leg runs on a column prefix and says so
        need = n_int * nc * K * 8 + 3 * n * nc
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        budget = 0.5 * host_memory_available() / max(local_world, 1)
        e_cols = nc if need <= budget else max(64, int(nc * budget / need))
        if os.environ.get("BNK_BENCH_E2E_COLS"):
            e_cols = min(nc, int(os.environ["BNK_BENCH_E2E_COLS"]))
        if world > 1:
            tcols = torch.tensor([e_cols], dtype=torch.int64, device=dev)
            dist_.all_reduce(tcols, op=dist_.ReduceOp.MIN)
            e_cols = int(tcols[0])
        e_obs = sh.obs_local if e_cols == nc else np.ascontiguousarray(sh.obs_local[:, :e_cols])
        e_pres = None if sh.present_local is None else np.ascontiguousarray(sh.present_local[:, :e_cols])
        e_rates = np.ascontiguousarray(sh.rates[:e_cols])
        h_obs = torch.from_numpy(e_obs).pin_memory()
        h_pres = None if e_pres is None else torch.from_numpy(e_pres).pin_memory()
        h_state = torch.empty((n, e_cols), dtype=torch.uint8).pin_memory()
        h_post = torch.empty((n_int, e_cols, K), dtype=torch.float64).pin_memory()
        h_one = torch.empty((1, e_cols, K), dtype=torch.float64).pin_memory()
        h_logl = torch.empty((e_cols,), dtype=torch.float64).pin_memory()
        ws2 = lib.Workspace(sh.model, sh.tree, e_cols, device=local_rank)   # AUTO evidence mode, private stream, pipelined
        L = lib.lib()
        one_q = np.array([int(sh.tree.internal_nodes()[n_int // 2])], dtype=np.int32)
        pp = None if h_pres is None else h_pres.data_ptr()

        def e_joint():
            ws2.set_distances(sh.dist0)
            ws2.set_rates(e_cols, e_rates)
            lib.check(L.bnk_joint(ws2.h, e_cols, h_obs.data_ptr(), pp, h_state.data_ptr()))

        def e_marg():
            ws2.set_distances(sh.dist0)
            lib.check(L.bnk_marginal(ws2.h, e_cols, h_obs.data_ptr(), pp, None, -1, h_post.data_ptr(), h_logl.data_ptr()))

        def e_both():  # both passes in one call: the joint pass of a chunk runs under the previous chunk's result copy
            ws2.set_distances(sh.dist0)
            ws2.set_rates(e_cols, e_rates)
            lib.check(L.bnk_joint_marginal(ws2.h, e_cols, h_obs.data_ptr(), pp, h_state.data_ptr(), None, -1, h_post.data_ptr(), h_logl.data_ptr()))

        def e_one():  # what a GRASP `-m N<k>` call costs: Prediction.getMarginal returns ONE ancestor (Prediction.java:555-597)
            ws2.set_distances(sh.dist0)
            lib.check(L.bnk_marginal(ws2.h, e_cols, h_obs.data_ptr(), pp, lib._i(one_q), 1, h_one.data_ptr(), h_logl.data_ptr()))

        def wall(fn, reps):
            fn()
            if world > 1:
                dist_.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            if world > 1:
                dist_.barrier()
            return (time.perf_counter() - t0) * 1e3 / reps
        k_e2e = max(1, min(args.steps, 3))
        j_ms, m_ms, o_ms, b_ms = wall(e_joint, k_e2e), wall(e_marg, k_e2e), wall(e_one, k_e2e), wall(e_both, k_e2e)
        # the box's pinned device->host rate, every rank copying at once (the ceiling of the marginal leg)
        d_tmp = sh.d_post.view(-1)[: min(sh.d_post.numel(), h_post.numel(), 1 << 27)]
        h_tmp = h_post.view(-1)[: d_tmp.numel()]

        def d2h():
            h_tmp.copy_(d_tmp, non_blocking=True)
        d2h_ms = wall(d2h, 3)
        d2h_gbs = d_tmp.numel() * 8 / 1e9 / (d2h_ms * 1e-3)
        if present is None:  # cheap invariants of what came back: posteriors sum to 1, states in range
            assert abs(float(h_post[0, 0].sum()) - 1.0) < 1e-9 and int(h_state.max()) < K
        e2e = dict(cols=e_cols, joint_ms=j_ms, marginal_ms=m_ms, one_ancestor_ms=o_ms, both_ms=b_ms, d2h_gbs=d2h_gbs,
                   h2d=n * e_cols * (2 if present is not None else 1),
                   d2h=n * e_cols + n_int * e_cols * K * 8 + e_cols * 8, need=need, launches=ws2.launch_count())
        ws2.close()
        del h_post, h_state, h_obs
    # ---- max over ranks
    e_vals = [e2e["joint_ms"], e2e["marginal_ms"], e2e["one_ancestor_ms"], -e2e["d2h_gbs"], e2e["both_ms"]] if e2e else [0.0, 0.0, 0.0, 0.0, 0.0]
    tmax, smax, ej, em, eo, ed, eb = max_over_ranks(torch, dist_, world, dev, [total_ms, serial_ms] + e_vals)

    main_cols = n_cols * (world if args.weak else 1)
    updates_step = 3.0 * n_int * main_cols
    ms_per_step = tmax / args.steps
    value = updates_step / (ms_per_step * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    upd = float(n_int) * sh.ncols   # rank 0's block: the kernels whose phase times are reported
    tbl = TABLE_BYTES * (n - 1) * spec["ncat"]
    # The fused down-sweep of the height-2 level (r2) never writes or reads the from-above vectors of the height-1
    # nodes it evaluates in the parent's thread: 2 x 160 B less per such (node, column) than SURVEY 8(d) counts.
    # The algorithmic bytes of the down-sweep are restated accordingly; the survey's own
    # figure stays in the line as `survey_GB`.
    fused = fused_height1_nodes(parent) if present is None else 0
    survey = {k: BYTES_PER_UPDATE[k] * upd + tbl for k in BYTES_PER_UPDATE}
    restated = {k: survey[k] - BYTES_SAVED_PER_DIRECT[k] * fused * sh.ncols for k in survey}
    kern = {"joint_up+traceback": (phase_ms[1] + phase_ms[2], restated["joint_up+traceback"]),
            "sum_up": (phase_ms[3], restated["sum_up"]),
            "sum_down": (phase_ms[4], restated["sum_down"])}
    dom = max(kern, key=lambda k: kern[k][0])
    per_kernel = {k: {"ms": round(v[0], 4), "algorithmic_GB": round(v[1] / 1e9, 3), "survey_GB": round(survey[k] / 1e9, 3),
                      "achieved_GBps": round(v[1] / 1e9 / (v[0] * 1e-3), 1) if v[0] > 0 else None,
                      "frac": round(v[1] / 1e9 / (v[0] * 1e-3) / peak, 4) if v[0] > 0 else None} for k, v in kern.items()}
    achieved = kern[dom][1] / 1e9 / (kern[dom][0] * 1e-3)
    traffic, traffic_src = None, None  # DRAM bytes of the dominant sweep per step, from this round's committed ncu capture
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        if tr.get("workload") == args.workload and sh.ncols == spec["cols"] and present is None:
            traffic = tr["bytes_per_step"].get(dom)
            traffic_src = tr.get("source")
            for k in per_kernel:   # what each sweep really moves through DRAM, over this run's event time
                b = tr["bytes_per_step"].get(k)
                if b and kern[k][0] > 0:
                    per_kernel[k]["dram_GB"] = round(b / 1e9, 3)
                    per_kernel[k]["dram_frac"] = round(b / 1e9 / (kern[k][0] * 1e-3) / peak, 4)
    except Exception:
        pass

    line = {"metric": "column x node partial updates/s", "value": value, "unit": "updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "gpu_launches": int(launches),
            "cols_per_gpu": sh.ncols,
            "step_api": "bnk_joint_dev + bnk_marginal_dev" if sh.serial else
                        "bnk_joint_marginal_dev (joint pass on a second stream beside the marginal pass)",
            "ms_per_step_serial": smax / args.steps,
            "phase_ms": {"tables": round(float(phase_ms[0]), 4), "joint_up": round(float(phase_ms[1]), 4),
                         "traceback": round(float(phase_ms[2]), 4), "sum_up": round(float(phase_ms[3]), 4),
                         "sum_down": round(float(phase_ms[4]), 4),
                         "of": "rank 0's block of %d columns; CUDA events of the serial pass (bnk_joint_dev, then "
                               "bnk_marginal_dev: %d more steps in the same run) -- the sweeps of the timed step overlap" % (sh.ncols, args.steps)},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "traffic_note": "DRAM bytes per step of the dominant sweep (all its launches): " + str(traffic_src),
                         "algorithmic_bytes": "SURVEY.md 8(d) per-update figures x %d updates + table bytes (`survey_GB`), restated for what "
                                              "the r2 kernels no longer move: minus %s B per (node, column) of the %d height-1 nodes under "
                                              "level-scheduled parents (message rows never expanded, from-above vectors never stored)"
                                              % (int(upd), json.dumps(BYTES_SAVED_PER_DIRECT), fused),
                         "per_kernel": per_kernel},
            "clocks": clocks}
    if parity is not None:
        line["parity"] = parity
    if e2e:
        scale = e2e["cols"] / max(sh.ncols, 1)
        both = eb
        line["e2e"] = {"value": updates_step * scale / (both * 1e-3), "unit": "updates/s", "ms_per_step": both,
                       "h2d_bytes_per_step": int(e2e["h2d"]), "d2h_bytes_per_step": int(e2e["d2h"]),
                       "api": "bnk_joint_marginal (joint states + posteriors of all ancestors) with pinned host buffers: column chunks "
                              "through two internal workspaces, each chunk's joint + marginal kernels under the previous chunk's result copy",
                       "ms_two_calls": ej + em,
                       "per_pass": {"joint": {"ms": ej, "updates_per_s": n_int * main_cols * scale / (ej * 1e-3),
                                              "d2h_bytes": int(n * e2e["cols"])},
                                    "marginal_all_ancestors": {"ms": em, "updates_per_s": 2.0 * n_int * main_cols * scale / (em * 1e-3),
                                                               "d2h_bytes": int(n_int * e2e["cols"] * K * 8)},
                                    "marginal_one_ancestor": {"ms": eo, "note": "Prediction.getMarginal(N<k>): one posterior row out"}},
                       "pinned_d2h_GBps_per_gpu_all_ranks_copying": -ed,
                       "marginal_floor_ms_at_that_rate": n_int * e2e["cols"] * K * 8 / 1e9 / (-ed) * 1e3,
                       "gpu_launches_e2e_leg": int(e2e["launches"])}
        if e2e["cols"] != sh.ncols:
            line["e2e"]["note"] = ("host memory limit: this leg ran on the first %d of %d columns per GPU "
                                   "(it pins %.1f GB per rank at full width)" % (e2e["cols"], sh.ncols, e2e["need"] / 1e9))
    sh.close()

    # ------------------------------------------------------------------ sub-results: the other BASELINE configurations
    subs = {}
    want_sub = [] if (args.no_sub or args.workload != "c3a" or args.weak or args.gaps > 0) else [s for s in args.sub.split(",") if s]
    sub_steps, sub_warm = max(3, min(args.steps, 5)), 3
    for name in want_sub:
        if name in ("c3b", "c4"):
            sspec = WORKLOADS[name]
            mname, sp, sd, sr, so = make_workload(sspec, rank)
            s2 = Shard(lib, torch, sspec, mname, sp, sd, sr, so, world, rank, local_rank)
            ph = np.zeros(5)

            def coll():
                s2.ws.sync()
                ph[:] += [s2.ws.phase_ms(i) for i in range(5)]
            ms_serial = None
            if not s2.serial:
                ms_serial = timed_steps(torch, dist_, world, s2.step_serial, 1, sub_steps, after_step=coll)
            ms = timed_steps(torch, dist_, world, s2.step, sub_warm, sub_steps, after_step=coll if s2.serial else None)
            if ms_serial is None:
                ms_serial = ms
            par = None
            if rank == 0 and not args.no_parity and s2.ncols > 0:
                par = parity_sample(torch, s2, mname, sp, sd, 8 if name == "c3b" else 6)
            ms, ms_serial = max_over_ranks(torch, dist_, world, dev, [ms, ms_serial])
            ni = s2.n_int
            u = 3.0 * ni * sspec["cols"]
            ph /= sub_steps
            tb = TABLE_BYTES * (s2.n - 1) * sspec["ncat"]
            algo = sum(BYTES_PER_UPDATE.values()) * ni * s2.ncols + 3 * tb - sum(BYTES_SAVED_PER_DIRECT.values()) * fused_height1_nodes(sp) * s2.ncols
            algo_all = algo   # rank 0's block in ms (max over ranks): the same quantity per rank
            subs[name] = {"workload": sspec["name"], "cols_per_gpu": s2.ncols, "steps": sub_steps, "warmup": sub_warm,
                          "ms_per_step": ms / sub_steps, "updates_per_s": u / (ms / sub_steps * 1e-3),
                          "ms_per_step_serial": ms_serial / sub_steps,
                          "roofline_frac_step": algo_all / 1e9 / (ms / sub_steps * 1e-3) / peak,
                          "phase_ms": {"tables": round(float(ph[0]), 4), "joint_up": round(float(ph[1]), 4),
                                       "traceback": round(float(ph[2]), 4), "sum_up": round(float(ph[3]), 4),
                                       "sum_down": round(float(ph[4]), 4)},
                          "roofline_frac_step_rank0": algo / 1e9 / (float(ph.sum()) * 1e-3) / peak if ph.sum() > 0 else None,
                          "parity": par}
            s2.close()
            del s2, so
        elif name == "c5":
            # the optimisation loop on C3a: per evaluation new rates (tables rebuilt), a log-likelihood sweep of the
            # rank's block, and the all-reduce of the per-rank sums -- all inside the timed region
            mname, sp, sd, sr, so = model_name, parent, dist, rates, obs
            s5 = Shard(lib, torch, spec, mname, sp, sd, sr, so, world, rank, local_rank, want_post=False)
            run = bd.ShardedRun(s5.ws, s5.total_cols, sr, s5.d_obs, None, s5.d_state, s5.d_logl, s5.d_sum)
            it = [0]
            sums = []

            def evaluate():
                it[0] += 1
                run.evaluate(rate_scale=1.0 + 0.01 * (it[0] % 7))
            evals = max(10, args.steps)
            ms = timed_steps(torch, dist_, world, evaluate, 3, evals)
            # the value the all-reduce carries, against the oracle on a sample of the rank's block is covered by
            # tests/test_gpu_golden.py::test_c5_likelihood_loop_matches_oracle; here: every rank holds the same sum
            total = float(run.sum.cpu()[0])
            sums.append(total)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            run.joint_and_gather(dst=0)   # untimed: the joint pass's buffers and the all_gather's channels are set up on first use
            if world > 1:
                dist_.barrier()
            torch.cuda.synchronize()
            g0.record()
            full = run.joint_and_gather(dst=0)
            g1.record()
            torch.cuda.synchronize()
            gather_ms = g0.elapsed_time(g1)
            ok_shape = (full is None) or (tuple(full.shape) == (s5.n, s5.total_cols))
            ms, gather_ms = max_over_ranks(torch, dist_, world, dev, [ms, gather_ms])
            same = max_over_ranks(torch, dist_, world, dev, [total, -total])
            subs["c5"] = {"workload": "C5 optimisation loop on C3a: rates perturbed per evaluation (tables rebuilt), logL sweep, "
                                      "all-reduce(sum) of the per-rank sums inside the timed region",
                          "cols_per_gpu": s5.ncols, "evaluations": evals, "ms_per_evaluation": ms / evals,
                          "updates_per_s": float(s5.n_int) * s5.total_cols / (ms / evals * 1e-3),
                          "collective": ("torch.distributed all_reduce (NCCL) of 1 x f64 per evaluation" if world > 1 else "none (one rank)"),
                          "sum_logl": total, "all_ranks_hold_the_same_sum": bool(same[0] == -same[1]),
                          "joint_plus_state_gather_ms": gather_ms,
                          "gathered_shape_ok": bool(ok_shape)}
            s5.close()
            del s5
        elif name == "indel":
            # section 8 f-3, the reference's real optimisation loop (IndelPeeler.optimiseMuLambda): per evaluation a new
            # GapSubstModel (host eigensystem 21 x 21), per-branch tables, the gap-augmented peel of the rank's column
            # block at C3a size with 30 % clade-wise gaps, and the all-reduce of the per-rank sums
            from bnkit_b200 import synth
            lo, hi = bd.shard_columns(obs.shape[1], world, rank)
            aln = np.ascontiguousarray(obs[:, lo:hi]).copy()
            leafm = np.ones(len(parent), bool)
            leafm[parent[parent >= 0]] = False
            grng = np.random.default_rng(11 + lo)
            size = np.ones(len(parent), np.int64)
            for v in range(len(parent) - 1, 0, -1):
                size[parent[v]] += size[v]
            for cidx in range(aln.shape[1]):   # whole clades lose the column until ~30 % of the extants are gaps
                gaps = 0
                while gaps < 0.3 * leafm.sum():
                    v = int(grng.integers(1, len(parent)))
                    if size[v] > len(parent) // 8:
                        continue
                    aln[v:v + size[v], cidx] = 255
                    gaps += (size[v] + 1) // 2
            tree_i = lib.Tree(parent, dist)
            hnd = lib.Indel(lib.Model(model_name), tree_i, max(aln.shape[1], 1))
            hnd.set_alignment(aln)
            geom = 1.0 / 3000.0
            red = torch.zeros(1, dtype=torch.float64, device=dev)
            it = [0]
            last = [0.0]

            def evaluate_indel():
                it[0] += 1
                hnd.set_rates(0.02 + 0.001 * (it[0] % 5), 0.02 + 0.001 * (it[0] % 5))
                _, ssum = hnd.column_logl(1.0, geom)
                red[0] = ssum
                if world > 1:
                    dist_.all_reduce(red, op=dist_.ReduceOp.SUM)
                last[0] = float(red[0])
            evals = max(5, min(args.steps, 10))
            l0 = hnd.launch_count()
            ms = timed_steps(torch, dist_, world, evaluate_indel, 3, evals)
            sweep_ms, tab_ms = hnd.phase_ms(1), hnd.phase_ms(0)
            (ms, sweep_ms) = max_over_ranks(torch, dist_, world, dev, [ms, sweep_ms])
            par = None
            if rank == 0 and not args.no_parity:
                from oracle import oracle as O
                O.build()
                selc = np.linspace(0, aln.shape[1] - 1, 12).astype(int)
                hnd.set_rates(0.02, 0.02)
                got, _ = hnd.column_logl(1.0, geom)
                want = O.indel_column_logl(O.GapModel(O.Model(model_name), 0.02, 0.02), parent, dist, np.ascontiguousarray(aln[:, selc]), 1.0, geom, nthreads=threads)
                par = {"cols": int(len(selc)), "logl_max_rel": float(np.max(np.abs(got[selc] - want) / np.abs(want))),
                       "against": "oracle/bnk_indel.c (IndelPeeler restated in log space), same inputs"}
            ni = int(tree_i.n_internal)
            # algorithmic bytes per (ancestor, column): the 21-entry message + two exponents + the all-gap flag, written
            # once and read once by the parent
            ibytes = 2.0 * (21 * 8 + 9) * ni * (aln.shape[1] + 1)
            subs["indel"] = {"workload": "IndelPeeler.optimiseMuLambda evaluations on C3a with 30 % clade-wise gaps: new GapSubstModel + tables + "
                                         "gap-augmented peel of all columns + all-reduce(sum) per evaluation",
                             "cols_per_gpu": int(aln.shape[1]), "evaluations": evals, "ms_per_evaluation": ms / evals,
                             "updates_per_s": float(ni) * obs.shape[1] / (ms / evals * 1e-3),
                             "phase_ms": {"tables": round(float(tab_ms), 4), "peel": round(float(sweep_ms), 4)},
                             "roofline_frac_peel_rank0": ibytes / 1e9 / (sweep_ms * 1e-3) / peak if sweep_ms > 0 else None,
                             "gpu_launches_per_evaluation": int((hnd.launch_count() - l0) // (evals + 3)),
                             "sum_logl": last[0], "parity": par}
            hnd.close()
            del hnd, aln

    if rank == 0:
        if subs:
            line["sub_results"] = subs
        if world == 1 and not args.no_cpu:
            v, sample, _ = cpu_oracle_rate(model_name, parent, dist, rates, obs, threads)
            line["cpu_baseline"] = {"value": v, "unit": "updates/s", "cores": threads, "kind": "port", "sample": sample,
                                    "jvm": java_probe()}
        emit(line)
    if world > 1:
        dist_.destroy_process_group()


if __name__ == "__main__":
    main()
