mkdir -p gpurun_out
./tools/microbench/layout_bw > gpurun_out/t7_layout_bw.txt 2>&1; cat gpurun_out/t7_layout_bw.txt
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu ) > gpurun_out/t7_pytest_a.log 2>&1; tail -n 4 gpurun_out/t7_pytest_a.log
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --sub c3b,c4 > gpurun_out/t7_bench.json 2> gpurun_out/t7_bench.err; tail -c 300 gpurun_out/t7_bench.err
python bench.py --gaps 0.3 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/t7_bench_c3a_gaps.json 2>> gpurun_out/t7_bench.err
for f in bench bench_c3a_gaps; do python -c "
import json,sys; d=json.load(open('gpurun_out/t7_%s.json' % sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['ms_per_step_serial'], d['phase_ms'], d.get('parity'))
for k,v in d.get('sub_results',{}).items(): print(k, v['ms_per_step'], v.get('ms_per_step_serial'), v['phase_ms'], v.get('parity'))" $f; done
