mkdir -p gpurun_out
BNK_DOWN2_MINB=2 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/t10_bench_minb2.json 2> gpurun_out/t10_bench.err
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-sub --no-parity > gpurun_out/t10_bench.json 2>> gpurun_out/t10_bench.err
for f in bench_minb2 bench; do python -c "
import json,sys; d=json.load(open('gpurun_out/t10_%s.json' % sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['ms_per_step_serial'], d['phase_ms'], d['roofline']['per_kernel'])" $f; done
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/t10_pytest_all.log 2>&1; tail -n 5 gpurun_out/t10_pytest_all.log
