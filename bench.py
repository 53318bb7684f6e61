#!/usr/bin/env python3
"""bench.py -- throughput of the GRASP per-column tree-inference path on B200.

One step = one pass of the hot path over one batch of synthetic columns:
    P(t)/log P(t) table build  +  joint reconstruction (max-product up-sweep + traceback)
    +  marginal reconstruction of ALL ancestors (sum-product up-sweep + down-sweep) + log-likelihoods
on BASELINE.json configs[2] (C3a): 10,000-leaf balanced tree x 5,000 columns per GPU, LG + Gamma4
(per-column category rates), FP64.  Metric = column x node partial updates / s, one update = one
(column, internal node) message computation in one sweep: 3 sweeps per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3a|c3b|small] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); columns are sharded, 5,000 per GPU (weak scaling),
no data-path collective; the only collective is the barrier / max-reduce of the timing.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes per update, SURVEY.md section 8(d) (store-all-partials schedule, FP64, K = 20)
BYTES_PER_UPDATE = {"joint_up+traceback": 361.0, "sum_up": 337.0, "sum_down": 832.0}
TABLE_BYTES = 3200.0  # per (branch, category) per sweep

WORKLOADS = {
    "c3a": dict(tree="balanced", leaves=10000, cols=5000, model="LG", ncat=4, name="C3a balanced 10k-leaf x 5k-col LG+G4"),
    "c3b": dict(tree="caterpillar", leaves=10000, cols=5000, model="LG", ncat=4, name="C3b caterpillar 10k-leaf x 5k-col LG+G4"),
    "small": dict(tree="balanced", leaves=512, cols=1024, model="LG", ncat=4, name="small 512-leaf x 1024-col LG+G4"),
    # BASELINE configs[3]: 100k-leaf random tree x 2k columns, WAG, one rate (use --no-e2e: 32 GB of posteriors)
    "c4": dict(tree="random", leaves=100000, cols=2000, model="WAG", ncat=1, name="C4 random 100k-leaf x 2k-col WAG"),
    # BASELINE configs[4]: repeated log-likelihood evaluations (rates perturbed every iteration => tables rebuilt)
    # with an all-reduce of the per-GPU sum of column log-likelihoods
    "c5": dict(tree="balanced", leaves=10000, cols=5000, model="LG", ncat=4, name="C5 logL loop on C3a, rates perturbed per evaluation, all-reduce of sum(logL)", loop=True),
}


def make_workload(spec, rank, cols=None):
    from bnkit_b200 import lib, synth
    leaves = spec["leaves"]
    n_cols = cols or spec["cols"]
    if spec["tree"] == "balanced":
        parent, dist = synth.balanced_tree(leaves, seed=1)
    elif spec["tree"] == "caterpillar":
        parent, dist = synth.caterpillar_tree(leaves, seed=2)
    else:
        parent, dist = synth.random_tree(leaves, seed=5)
    model = lib.Model(spec["model"])
    cat_rates = synth.gamma_category_rates(0.5, spec["ncat"]) if spec["ncat"] > 1 else np.ones(1)
    rng = np.random.default_rng(4 + 1000 * rank)
    rates = cat_rates[rng.integers(0, spec["ncat"], n_cols)]
    if leaves <= 20000:
        obs = synth.simulate_alignment(parent, dist, model.parts(), n_cols, rates, seed=3 + 1000 * rank)
    else:  # very large trees: a conserved base state per leaf with 20 % noise (numpy simulation would take minutes)
        leaf = np.ones(len(parent), dtype=bool)
        leaf[parent[parent >= 0]] = False
        nl = int(leaf.sum())
        base = rng.integers(0, model.K, (nl, 1)).astype(np.uint8)
        noise = rng.integers(0, model.K, (nl, n_cols)).astype(np.uint8)
        obs = np.full((len(parent), n_cols), 255, dtype=np.uint8)
        obs[leaf] = np.where(rng.random((nl, n_cols)) < 0.2, noise, base)
    return model, parent, dist, rates, obs


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


def cpu_oracle_rate(spec, model_name, parent, dist, rates, obs, budget_s, threads):
    """Times the CPU restatement (oracle/) on a bounded column sample; returns (updates/s, sample text)."""
    from oracle import oracle as O
    O.build()
    om = O.Model(model_name)
    n_int = int((np.bincount(parent[parent >= 0], minlength=len(parent)) > 0).sum())
    def run(nc):
        o = np.ascontiguousarray(obs[:, :nc])
        r = np.ascontiguousarray(rates[:nc])
        t0 = time.perf_counter()
        O.fast_joint(om, parent, dist, o, None, r, nthreads=threads)
        O.fast_marginal(om, parent, dist, o, None, r, nthreads=threads, want_post=True)
        return time.perf_counter() - t0
    nc = min(obs.shape[1], threads)
    run(nc)                       # warm (page in the library)
    t = run(nc)
    for _ in range(4):            # grow the sample until it costs about budget_s (the per-rate table build is a fixed cost)
        if t >= 0.7 * budget_s or nc >= obs.shape[1]:
            break
        grow = min(4.0, max(1.5, budget_s / max(t, 1e-3)))
        nc = min(obs.shape[1], max(nc + threads, int(nc * grow) // threads * threads))
        t = run(nc)
    return 3.0 * n_int * nc / t, "%d of %d columns, all %d nodes, joint + marginal(all ancestors), %d threads, %.1f s" % (
        nc, obs.shape[1], len(parent), threads, t), t


def bind_to_gpu_numa_node(local_rank):
    """Run this rank (and first-touch its pinned buffers) on the NUMA node its GPU hangs off, so that the 8 GB result
    copies of several ranks do not cross sockets.  Best effort: any missing piece leaves the affinity untouched."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3a", choices=sorted(WORKLOADS))
    ap.add_argument("--cols", type=int, default=0, help="columns per GPU (default: the workload's)")
    ap.add_argument("--tile-cols", type=int, default=0)
    ap.add_argument("--cpt", type=int, default=0)
    ap.add_argument("--gaps", type=float, default=0.0, help="gap fraction at the leaves; > 0 adds a per-column presence mask (general kernels)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
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
    spec = WORKLOADS[args.workload]
    n_cols = args.cols or spec["cols"]
    threads = os.cpu_count() or 1
    config = {"workload": spec["name"], "tree": spec["tree"], "leaves": spec["leaves"], "cols_per_gpu": n_cols,
              "model": spec["model"] + "+G%d" % spec["ncat"], "passes": "tables + joint(up+traceback) + marginal(up+down, all ancestors)",
              "sharding": "columns, %d per GPU" % n_cols,
              "schedule": "wide height levels as level launches (one column per thread, messages through HBM) + column-owned walker for the narrow top",
              "l2": "per-step working set (messages 8 GB + posteriors 8 GB + pointers 2 GB at C3) >> 126 MB L2; no flush needed"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        from bnkit_b200 import synth  # workload generator only (numpy)
        model, parent, dist, rates, obs = make_workload(spec, 0, min(n_cols, 4 * threads * 8))
        per_step = max(2.0, min(20.0, 120.0 / max(args.steps + args.warmup, 1)))
        vals, sample = [], ""
        for i in range(args.warmup + args.steps):
            v, sample, _ = cpu_oracle_rate(spec, spec["model"], parent, dist, rates, obs, per_step, threads)
            if i >= args.warmup:
                vals.append(v)
        value = float(np.mean(vals))
        n_int = int((np.bincount(parent[parent >= 0], minlength=len(parent)) > 0).sum())
        line = {"impl": "reference", "metric": "column x node partial updates/s", "value": value, "unit": "updates/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * 3.0 * n_int * n_cols / value, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": "updates/s", "cores": threads, "kind": "port", "sample": sample,
                                 "note": "C restatement of bnkit's VarElim path (oracle/); the JVM reference cannot run here (no JDK)"},
                "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist_
    from bnkit_b200 import lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libbnkgpu has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    model, parent, dist, rates, obs = make_workload(spec, rank, n_cols)
    tree = lib.Tree(parent, dist)
    n, n_int, K = tree.n, tree.n_internal, model.K
    stream = torch.cuda.current_stream()
    ws = lib.Workspace(model, tree, n_cols, device=local_rank, stream=stream.cuda_stream)
    if args.tile_cols or args.cpt:
        ws.configure(args.tile_cols, args.cpt)

    d_present = None
    if args.gaps > 0:  # per-column forests: gaps at the leaves, ancestors present on paths between present leaves
        from bnkit_b200 import synth
        g_rng = np.random.default_rng(77 + rank)
        leaf = np.ones(n, dtype=bool)
        leaf[parent[parent >= 0]] = False
        obs = obs.copy()
        obs[(g_rng.random(obs.shape) < args.gaps) & leaf[:, None]] = 255
        present = synth.path_presence_mask(parent, (obs != 255) & leaf[:, None])
        d_present = torch.from_numpy(present).to(dev)
        config["gaps"] = args.gaps
        args.no_e2e = True
        args.no_cpu = True
    d_obs = torch.from_numpy(obs).to(dev)
    d_state = torch.empty((n, n_cols), dtype=torch.uint8, device=dev)
    d_post = torch.empty((1 if spec.get("loop") else n_int, n_cols, K), dtype=torch.float64, device=dev)
    d_logl = torch.empty((n_cols,), dtype=torch.float64, device=dev)

    loop_mode = bool(spec.get("loop"))
    d_sum = torch.zeros((1,), dtype=torch.float64, device=dev)
    it_count = [0]

    def step_resident():
        if loop_mode:  # one log-likelihood evaluation of an optimisation loop
            it_count[0] += 1
            ws.set_rates(n_cols, rates * (1.0 + 0.01 * (it_count[0] % 7)))
            ws.loglik_dev(n_cols, d_obs, None, d_logl, d_sum)
            if world > 1:
                dist_.all_reduce(d_sum, op=dist_.ReduceOp.SUM)  # the only data-path collective: one double
            return
        ws.set_rates(n_cols, rates)  # marks the tables dirty: every step rebuilds P(t) / log P(t)
        ws.joint_dev(n_cols, d_obs, d_present, d_state)
        ws.marginal_dev(n_cols, d_obs, d_present, d_post, d_logl)

    def barrier():
        if world > 1:
            dist_.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    phase_ms = np.zeros(5)
    launches0 = ws.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
        ws.sync()
        phase_ms += [ws.phase_ms(i) for i in range(5)]
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    launches = ws.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    phase_ms /= args.steps

    # ---- end to end through the host-buffer ABI (pinned host memory, copies inside the timed region)
    e2e = None
    if loop_mode:
        args.no_e2e = True
        args.no_cpu = True
    e2e_cols = n_cols
    if not args.no_e2e:
        # The leg pins n_int x n_cols x K doubles of host memory per rank (8 GB at C3).  On a box whose host memory
        # (or cgroup limit) cannot hold that for every local rank, the leg runs on a column prefix instead and says
        # so: the rate is D2H-bound per column either way.  Rank 0 decides for everybody.
        need = n_int * n_cols * K * 8 + 3 * n * n_cols
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        avail = host_memory_available()
        budget = 0.5 * avail / max(local_world, 1)
        if need > budget:
            e2e_cols = max(64, min(n_cols, int(n_cols * budget / need)))
        if os.environ.get("BNK_BENCH_E2E_COLS"):
            e2e_cols = min(n_cols, int(os.environ["BNK_BENCH_E2E_COLS"]))
        if world > 1:
            tcols = torch.tensor([e2e_cols], dtype=torch.int64, device=dev)
            dist_.broadcast(tcols, src=0)
            e2e_cols = int(tcols[0])
        e_obs = obs if e2e_cols == n_cols else np.ascontiguousarray(obs[:, :e2e_cols])
        e_rates = rates if (rates is None or e2e_cols == n_cols) else np.ascontiguousarray(rates[:e2e_cols])
        h_obs = torch.from_numpy(e_obs).pin_memory()
        h_state = torch.empty((n, e2e_cols), dtype=torch.uint8).pin_memory()
        h_post = torch.empty((n_int, e2e_cols, K), dtype=torch.float64).pin_memory()
        h_logl = np.empty(e2e_cols)
        def step_e2e():
            ws.set_rates(e2e_cols, e_rates)
            lib.check(lib.lib().bnk_joint(ws.h, e2e_cols, h_obs.data_ptr(), None, h_state.data_ptr()))
            lib.check(lib.lib().bnk_marginal(ws.h, e2e_cols, h_obs.data_ptr(), None, None, -1, h_post.data_ptr(),
                                             h_logl.ctypes.data))
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        k_e2e = max(1, min(args.steps, 3))
        for _ in range(k_e2e):
            step_e2e()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / k_e2e
        h2d = 2 * n * e2e_cols
        d2h = n * e2e_cols + n_int * e2e_cols * K * 8 + e2e_cols * 8
        # parity of what came back (cheap invariants): posteriors sum to 1, states in range
        assert abs(float(h_post[0, 0].sum()) - 1.0) < 1e-9 and int(h_state.max()) < K
    # ---- max over ranks
    tmax, e2emax = total_ms, (e2e_ms if not args.no_e2e else 0.0)
    if world > 1:
        tt = torch.tensor([total_ms, e2emax], dtype=torch.float64, device=dev)
        dist_.all_reduce(tt, op=dist_.ReduceOp.MAX)
        tmax, e2emax = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            dist_.destroy_process_group()
        return

    updates_step = (1.0 if loop_mode else 3.0) * n_int * n_cols * world
    ms_per_step = tmax / args.steps
    value = updates_step / (ms_per_step * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    upd = float(n_int) * n_cols
    tbl = TABLE_BYTES * (n - 1) * spec["ncat"]
    kern = {"joint_up+traceback": (phase_ms[1] + phase_ms[2], BYTES_PER_UPDATE["joint_up+traceback"] * upd + tbl),
            "sum_up": (phase_ms[3], BYTES_PER_UPDATE["sum_up"] * upd + tbl),
            "sum_down": (phase_ms[4], BYTES_PER_UPDATE["sum_down"] * upd + tbl)}
    dom = max(kern, key=lambda k: kern[k][0])
    per_kernel = {k: {"ms": round(v[0], 4), "algorithmic_GB": round(v[1] / 1e9, 3),
                      "achieved_GBps": round(v[1] / 1e9 / (v[0] * 1e-3), 1) if v[0] > 0 else None,
                      "frac": round(v[1] / 1e9 / (v[0] * 1e-3) / peak, 4) if v[0] > 0 else None} for k, v in kern.items()}
    achieved = kern[dom][1] / 1e9 / (kern[dom][0] * 1e-3)
    traffic = None  # DRAM bytes of the dominant sweep per step, from the committed ncu capture of this workload
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        if tr.get("workload") == args.workload and n_cols == spec["cols"]:
            traffic = tr["bytes_per_step"].get(dom)
    except Exception:
        pass
    line = {"metric": "column x node partial updates/s", "value": value, "unit": "updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "gpu_launches": int(launches),
            "phase_ms": {"tables": round(float(phase_ms[0]), 4), "joint_up": round(float(phase_ms[1]), 4),
                         "traceback": round(float(phase_ms[2]), 4), "sum_up": round(float(phase_ms[3]), 4),
                         "sum_down": round(float(phase_ms[4]), 4)},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "traffic_note": "DRAM bytes per step of the dominant sweep (all its launches), profiles/roofline_traffic.json",
                         "algorithmic_bytes": "SURVEY.md 8(d) per-update figures x %d updates + table bytes" % int(upd),
                         "per_kernel": per_kernel},
            "clocks": clocks}
    if not args.no_e2e:
        line["e2e"] = {"value": updates_step * (e2e_cols / n_cols) / (e2emax * 1e-3), "unit": "updates/s", "ms_per_step": e2emax,
                       "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}
        if numa_node is not None:
            line["e2e"]["numa"] = "ranks bound to their GPU's NUMA node"
        if e2e_cols != n_cols:
            line["e2e"]["note"] = ("host memory limit: this leg ran on the first %d of %d columns per GPU "
                                   "(it pins %.1f GB per rank at full width)" % (e2e_cols, n_cols, need / 1e9))
    if world == 1 and not args.no_cpu:
        v, sample, _ = cpu_oracle_rate(spec, spec["model"], parent, dist, rates, obs, args.cpu_budget, threads)
        line["cpu_baseline"] = {"value": v, "unit": "updates/s", "cores": threads, "kind": "port", "sample": sample}
    if rank == 0:
        emit(line)
    if world > 1:
        dist_.destroy_process_group()


if __name__ == "__main__":
    main()
