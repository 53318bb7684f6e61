mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu ) > gpurun_out/t16_pytest.log 2>&1; tail -n 3 gpurun_out/t16_pytest.log
for v in a b; do python bench.py --no-e2e --no-cpu --sub c4 --steps 10 > gpurun_out/t16_bench_$v.json 2>> gpurun_out/t16.err; python -c "
import json; d=json.load(open('gpurun_out/t16_bench_$v.json')); print('$v', d['ms_per_step'], d['phase_ms']['tables'], d['sub_results']['c4']['phase_ms']['tables'], d['parity']['joint_equal'], d['parity']['post_max_rel'])"; done
tail -c 300 gpurun_out/t16.err
