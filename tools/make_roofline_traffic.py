#!/usr/bin/env python3
"""profiles/roofline_traffic.json from an ncu traffic capture of one bench.py step:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file traffic.csv \
        python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-sub
    python tools/make_roofline_traffic.py traffic.csv <tag>

Kernels are assigned to the sweep whose phase timer covers them (bench.py's roofline.per_kernel): the joint up-sweep
with its traceback, the sum-product up-sweep, the down-sweep.  bench.py reads the file for `roofline.traffic`.
"""
import csv, json, shutil, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
src, tag = Path(sys.argv[1]), sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}


def sweep_of(name):
    n = name.replace(" ", "")
    if any(k in n for k in ("cherry_pairs_kernel<20,0>", "cherry_expand_kernel<20,0>", "up_wide_kernel<20,0>", "up_warp_kernel<0", "up_wide_gen_kernel<20,0>",
                            "traceback", "leaf_states_kernel", "scatter_rows_kernel")):
        return "joint_up+traceback"
    if any(k in n for k in ("cherry_pairs_kernel<20,1>", "cherry_expand_kernel<20,1>", "up_wide_kernel<20,1>", "up_mma_kernel", "up_warp_kernel<1",
                            "up_wide_gen_kernel<20,1>", "sum_escale_kernel")):
        return "sum_up"
    if any(k in n for k in ("down_wide", "down_warp_kernel")):
        return "sum_down"
    return None


tot, launches, seen = {}, {}, {}
for r in rows[1:]:
    s = sweep_of(r[ix["Kernel Name"]])
    if s is None:
        continue
    tot[s] = tot.get(s, 0) + int(float(r[ix["Metric Value"]].replace(",", "")))
    seen.setdefault(s, set()).add(r[ix["ID"]])
for s in seen:
    launches[s] = len(seen[s])
out = {"source": "profiles/%s_traffic.csv: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one bench.py step on C3a "
                 "(balanced 10k leaves x 5k columns, LG+G4), summed over the launches of each sweep by tools/make_roofline_traffic.py" % tag,
       "workload": "c3a", "bytes_per_step": tot, "launches": launches}
dst = ROOT / "profiles" / ("%s_traffic.csv" % tag)
if src.resolve() != dst.resolve():
    shutil.copy(src, dst)
(ROOT / "profiles" / "roofline_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out))
