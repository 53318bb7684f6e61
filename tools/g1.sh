mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu ) > gpurun_out/t1_pytest_a.log 2>&1; tail -5 gpurun_out/t1_pytest_a.log
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "c3a_full or c3b_full or c3b_masked" ) > gpurun_out/t1_pytest_b.log 2>&1; tail -5 gpurun_out/t1_pytest_b.log
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --sub c3b,c4 > gpurun_out/t1_bench.json 2> gpurun_out/t1_bench.err; tail -c 300 gpurun_out/t1_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/t1_bench.json'))
print('c3a', d['ms_per_step'], d['ms_per_step_serial'], d['phase_ms'])
for k,v in d.get('sub_results',{}).items(): print(k, v['ms_per_step'], v.get('ms_per_step_serial'), v['phase_ms'], v.get('parity'))
print(d.get('parity'))
PY
python bench.py --workload c3b --gaps 0.3 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/t1_bench_c3b_gaps.json 2>> gpurun_out/t1_bench.err
python -c "
import json; d=json.load(open('gpurun_out/t1_bench_c3b_gaps.json')); print('c3b gaps', d['ms_per_step'], d['ms_per_step_serial'], d['phase_ms'], d.get('parity'))"
ncu --set full --clock-control none --import-source on -k regex:"down_wide_kernel" -c 10 -o gpurun_out/prof_t1_down python bench.py --serial --steps 1 --warmup 0 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/t1_ncu.log 2>&1
ncu -i gpurun_out/prof_t1_down.ncu-rep --page raw --csv > gpurun_out/prof_t1_down_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_t1_down.ncu-rep --page source --csv --print-kernel-base function -k regex:"down_wide_kernel<20, *false>|down_wide_kernel<20, 0>" 2>/dev/null | head -c 3000000 > gpurun_out/prof_t1_down_source.csv
rm -f gpurun_out/prof_t1_down.ncu-rep
ls -la gpurun_out | tail -12
