#!/bin/bash
# One GPU session producing the artefacts profiles/ is built from (tag = $1).
tag=${1:-r2_w1}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -n 3 gpurun_out/pytest_gpu_$tag.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py > gpurun_out/bench_c3a_$tag.json 2>> gpurun_out/bench_$tag.err; tail -c 600 gpurun_out/bench_c3a_$tag.json; echo
python bench.py --workload c3b --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c3b_$tag.json 2>> gpurun_out/bench_$tag.err
python bench.py --gaps 0.3 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c3a_gaps_$tag.json 2>> gpurun_out/bench_$tag.err
python bench.py --workload c3b --gaps 0.3 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c3b_gaps_$tag.json 2>> gpurun_out/bench_$tag.err
python tools/bench_bep.py --cpu-budget 5 > gpurun_out/bep_c3_$tag.json 2>> gpurun_out/bench_$tag.err
python tools/bench_real.py > gpurun_out/bench_real_$tag.json 2>> gpurun_out/bench_$tag.err
python tools/bench_indel.py --brent > gpurun_out/indel_c3_$tag.json 2>> gpurun_out/bench_$tag.err
for f in c3b c3a_gaps c3b_gaps; do python -c "import json,sys; d=json.load(open('gpurun_out/bench_${f}_$tag.json')); print('$f', d['ms_per_step'], d['ms_per_step_serial'], d['phase_ms'])"; done
python -c "import json; d=json.load(open('gpurun_out/bep_c3_$tag.json')); print('bep', d['seconds'], d['info'])"
cat gpurun_out/bench_real_$tag.json gpurun_out/indel_c3_$tag.json
# launch list (serialised, cold cache: shares only) and DRAM traffic per launch; --serial: one pass after the other, one loop
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --serial --steps 2 --warmup 1 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/ncu1_$tag.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 260 --csv --log-file gpurun_out/traffic_$tag.csv python bench.py --serial --steps 1 --warmup 0 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/ncu2_$tag.log 2>&1
# full captures: the level kernels of the default workload (up-sweeps; fused down-sweep), the walkers on the caterpillar
# (unmasked, masked), the indel and BEP kernels
ncu --set full --clock-control none --import-source on -k regex:"up_wide_kernel|cherry_pairs" -c 8 -o gpurun_out/prof_${tag}_up python bench.py --serial --steps 1 --warmup 0 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/ncu3_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"down_wide2|traceback_cherry" -c 10 -o gpurun_out/prof_${tag}_down python bench.py --serial --steps 1 --warmup 0 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/ncu3b_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"warp_kernel|mma_kernel|traceback_cols" -c 4 -o gpurun_out/prof_${tag}_walker python bench.py --workload c3b --serial --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity > gpurun_out/ncu4_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"warp_kernel|walk_case|traceback_cols" -c 5 -o gpurun_out/prof_${tag}_walker_masked python bench.py --workload c3b --gaps 0.3 --serial --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity > gpurun_out/ncu4b_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"indel" -c 14 -o gpurun_out/prof_${tag}_indel python tools/bench_indel.py --evals 1 --parity-cols 0 > gpurun_out/ncu5_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bep_reach|bep_forward_level|bep_backward_level" -c 8 -o gpurun_out/prof_${tag}_bep python tools/bench_bep.py --cpu-budget 0 > gpurun_out/ncu6_$tag.log 2>&1
# gpurun merges at most 64 MiB back: keep the raw-metric pages as CSV and drop the reports
for r in up down walker walker_masked indel bep; do
  ncu -i gpurun_out/prof_${tag}_$r.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_${r}_raw.csv 2>/dev/null
  rm -f gpurun_out/prof_${tag}_$r.ncu-rep
done
ls -la gpurun_out/*$tag* | awk '{print $5, $9}'
