mkdir -p gpurun_out
./tools/microbench/layout_bw > gpurun_out/t4_layout_bw.txt 2>&1; cat gpurun_out/t4_layout_bw.txt
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu ) > gpurun_out/t4_pytest_a.log 2>&1; tail -n 5 gpurun_out/t4_pytest_a.log
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "c3b_full or c3b_masked or c3a_masked" ) > gpurun_out/t4_pytest_b.log 2>&1; tail -n 5 gpurun_out/t4_pytest_b.log
python bench.py --workload c3b --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/t4_bench_c3b.json 2> gpurun_out/t4_bench.err; tail -c 300 gpurun_out/t4_bench.err
python bench.py --workload c3b --gaps 0.3 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/t4_bench_c3b_gaps.json 2>> gpurun_out/t4_bench.err
python bench.py --gaps 0.3 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/t4_bench_c3a_gaps.json 2>> gpurun_out/t4_bench.err
BNK_WALKER=lane python bench.py --workload c3b --gaps 0.3 --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/t4_bench_c3b_gaps_lane.json 2>> gpurun_out/t4_bench.err
for f in c3b c3b_gaps c3a_gaps c3b_gaps_lane; do python -c "
import json,sys; d=json.load(open('gpurun_out/t4_bench_%s.json' % sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['ms_per_step_serial'], d['phase_ms'], d.get('parity'))" $f; done
