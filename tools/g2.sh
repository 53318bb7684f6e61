mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_height2 or large_irregular or cherry or uninstantiated or caterpillar or masked_walker" ) > gpurun_out/t2_pytest_a.log 2>&1; tail -5 gpurun_out/t2_pytest_a.log
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "c3a_full or c4_full or c3b_full" ) > gpurun_out/t2_pytest_b.log 2>&1; tail -5 gpurun_out/t2_pytest_b.log
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --sub c3b,c4 > gpurun_out/t2_bench.json 2> gpurun_out/t2_bench.err; tail -c 300 gpurun_out/t2_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/t2_bench.json'))
print('c3a', d['ms_per_step'], d['ms_per_step_serial'], d['phase_ms'])
for k,v in d.get('sub_results',{}).items(): print(k, v['ms_per_step'], v.get('ms_per_step_serial'), v['phase_ms'], v.get('parity'))
print(d.get('parity'))
PY
BNK_NO_FUSE2=1 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/t2_bench_nofuse.json 2>> gpurun_out/t2_bench.err
python -c "
import json; d=json.load(open('gpurun_out/t2_bench_nofuse.json')); print('c3a nofuse2', d['ms_per_step'], d['ms_per_step_serial'], d['phase_ms'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/t2_launches_c3b.csv python bench.py --workload c3b --serial --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/t2_ncu1.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/t2_launches_c3b.csv')) if len(r)>10 and r[0]!='ID']
for r in rows[-24:]: print(r[4][:70], r[8], r[-1])
PY
ncu --set full --clock-control none --import-source on -k regex:"traceback_cols|down_wide2" -c 3 -o gpurun_out/prof_t2 python bench.py --workload c3b --serial --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity > gpurun_out/t2_ncu2.log 2>&1
ncu -i gpurun_out/prof_t2.ncu-rep --page raw --csv > gpurun_out/prof_t2_raw.csv 2>/dev/null
rm -f gpurun_out/prof_t2.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"down_wide2" -c 1 -o gpurun_out/prof_t2b python bench.py --serial --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity --no-sub > gpurun_out/t2_ncu3.log 2>&1
ncu -i gpurun_out/prof_t2b.ncu-rep --page raw --csv > gpurun_out/prof_t2b_raw.csv 2>/dev/null
rm -f gpurun_out/prof_t2b.ncu-rep
