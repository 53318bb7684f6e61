mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu ) > gpurun_out/t5_pytest_a.log 2>&1; tail -n 4 gpurun_out/t5_pytest_a.log
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "c3a_masked or c3b_masked" ) > gpurun_out/t5_pytest_b.log 2>&1; tail -n 4 gpurun_out/t5_pytest_b.log
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/t5_bench.json 2> gpurun_out/t5_bench.err; tail -c 300 gpurun_out/t5_bench.err
BNK_DOWN_EARLY=1 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/t5_bench_early.json 2>> gpurun_out/t5_bench.err
python bench.py --workload c3b --gaps 0.3 --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/t5_bench_c3b_gaps.json 2>> gpurun_out/t5_bench.err
python bench.py --workload c3b --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/t5_bench_c3b.json 2>> gpurun_out/t5_bench.err
for f in bench bench_early bench_c3b_gaps bench_c3b; do python -c "
import json,sys; d=json.load(open('gpurun_out/t5_%s.json' % sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['ms_per_step_serial'], d['phase_ms'])" $f; done
