#!/usr/bin/env python3
"""Summarise gpurun_out ncu artefacts into profiles/ (tracked).

    python tools/summarize_profiles.py <tag> <launches.csv> <bench.json> <report.ncu-rep | raw.csv> [more reports ...]
"""
import csv, json, shutil, subprocess, sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag, launches, bench = sys.argv[1], Path(sys.argv[2]), Path(sys.argv[3])
reps = [Path(x) for x in sys.argv[4:]]
out = ROOT / "profiles"
out.mkdir(exist_ok=True)

lines = ["# ncu summary %s" % tag, ""]
if bench and bench.exists():
    b = json.loads(bench.read_text().strip().splitlines()[-1])
    shutil.copy(bench, out / ("%s_bench.json" % tag))
    lines += ["Bench line (CUDA events, no profiler): %.3f ms/step, %.4g updates/s, phases %s, e2e %s" % (
        b["ms_per_step"], b["value"], json.dumps(b["phase_ms"]), json.dumps(b.get("e2e"))), ""]

rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
keep = [hdr] + rows[1:]
with open(out / ("%s_launches.csv" % tag), "w", newline="") as fh:
    csv.writer(fh).writerows(keep)
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    k = r[idx["Kernel Name"]].split("(")[0]
    v = float(r[idx["Metric Value"]].replace(",", "")) / 1e6
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; serialised, cold cache: compare shares)", "",
          "| kernel | launches | total ms | ms / launch | share |", "|---|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append("| `%s` | %d | %.3f | %.3f | %.3f |" % (k, v[0], v[1], v[1] / v[0], v[1] / tot))

recs = []   # one dict per profiled launch: metric name -> (value, unit); every report carries its own unit row
for rep in reps:
    if rep.suffix == ".csv":   # the raw page exported on the GPU box (tools/gpu_round.sh keeps these: the reports exceed gpurun's 64 MiB)
        raw = rep.read_text()
    else:
        raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    part = list(csv.reader(raw.splitlines()))
    if len(part) < 3:
        continue
    for r in part[2:]:
        recs.append({x: (r[i], part[1][i]) for i, x in enumerate(part[0]) if i < len(r)})
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (occupancy)"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__block_size", "block"), ("launch__grid_size", "grid"),
        ("smsp__inst_executed.sum", "warp instructions"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"), ("lts__t_bytes.sum", "L2 bytes")]
lines += ["", "## `ncu --set full --clock-control none --import-source on` (one launch each)", ""]
for rec in recs:
    lines.append("### `%s`  grid %s" % (rec["Kernel Name"][0], rec.get("launch__grid_size", ("?", ""))[0]))
    lines.append("")
    for key, name in want:
        if key in rec:
            lines.append("- %s: %s %s" % (name, rec[key][0], rec[key][1]))
    stalls = [x for x in rec if "issue_stalled" in x and x.endswith("_per_issue_active.ratio")]
    top = sorted([(float(rec[s][0]), s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                  for s in stalls if rec[s][0]], reverse=True)[:6]
    lines.append("- warp stalls per issued instruction: " + ", ".join("%s %.2f" % (n, v) for v, n in top))
    if "dram__bytes_read.sum" in rec:
        def gb(key):
            v, u = float(rec[key][0]), rec[key][1]
            return v * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}.get(u, 1.0)
        lines.append("- DRAM traffic per launch: %.3f GB" % (gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")))
    lines.append("")
(out / ("%s_ncu_summary.md" % tag)).write_text("\n".join(lines) + "\n")
print("wrote", out / ("%s_ncu_summary.md" % tag))
