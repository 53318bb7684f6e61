mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu ) > gpurun_out/t3_pytest_a.log 2>&1; tail -n 5 gpurun_out/t3_pytest_a.log
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "c3a_full or c3b_full or c3b_masked" ) > gpurun_out/t3_pytest_b.log 2>&1; tail -n 5 gpurun_out/t3_pytest_b.log
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --sub c3b,c4 > gpurun_out/t3_bench.json 2> gpurun_out/t3_bench.err; tail -c 300 gpurun_out/t3_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/t3_bench.json'))
print('c3a', d['ms_per_step'], d['ms_per_step_serial'], d['phase_ms'])
for k,v in d.get('sub_results',{}).items(): print(k, v['ms_per_step'], v.get('ms_per_step_serial'), v['phase_ms'], v.get('parity'))
print(d.get('parity'))
PY
BNK_DOWN_PIPE=0 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/t3_bench_nopipe.json 2>> gpurun_out/t3_bench.err
python -c "
import json; d=json.load(open('gpurun_out/t3_bench_nopipe.json')); print('c3a nopipe', d['ms_per_step'], d['ms_per_step_serial'], d['phase_ms'])"
for cfg in "BNK_WARP_C=2" "BNK_WARP_C=2 BNK_WARP_W=4" "BNK_WARP_W=5" "BNK_MMA_W=4"; do
  env $cfg python bench.py --workload c3b --steps 3 --warmup 2 --no-e2e --no-cpu --no-parity > gpurun_out/t3_c3b_ab.json 2>> gpurun_out/t3_bench.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/t3_c3b_ab.json')); print('c3b', sys.argv[1:], d['ms_per_step'], d['ms_per_step_serial'], d['phase_ms'])" $cfg
done
ncu --set full --clock-control none --import-source on -k regex:"down_wide_pipe" -c 8 -o gpurun_out/prof_t3 python bench.py --serial --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity --no-sub > gpurun_out/t3_ncu2.log 2>&1
ncu -i gpurun_out/prof_t3.ncu-rep --page raw --csv > gpurun_out/prof_t3_raw.csv 2>/dev/null
rm -f gpurun_out/prof_t3.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"traceback_cols" -c 1 -o gpurun_out/prof_t3b python bench.py --workload c3b --serial --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity > gpurun_out/t3_ncu3.log 2>&1
ncu -i gpurun_out/prof_t3b.ncu-rep --page raw --csv > gpurun_out/prof_t3b_raw.csv 2>/dev/null
rm -f gpurun_out/prof_t3b.ncu-rep
