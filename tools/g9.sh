mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu ) > gpurun_out/t9_pytest_a.log 2>&1; tail -n 4 gpurun_out/t9_pytest_a.log
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "c3a_full or c4_full" ) > gpurun_out/t9_pytest_b.log 2>&1; tail -n 4 gpurun_out/t9_pytest_b.log
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --sub c4 > gpurun_out/t9_bench.json 2> gpurun_out/t9_bench.err; tail -c 300 gpurun_out/t9_bench.err
BNK_CHERRY_DIRECT=0 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/t9_bench_nodirect.json 2>> gpurun_out/t9_bench.err
for f in bench bench_nodirect; do python -c "
import json,sys; d=json.load(open('gpurun_out/t9_%s.json' % sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['ms_per_step_serial'], d['phase_ms'], d.get('parity'))
for k,v in d.get('sub_results',{}).items(): print(k, v['ms_per_step'], v.get('ms_per_step_serial'), v['phase_ms'], v.get('parity'))" $f; done
