mkdir -p gpurun_out
( BNK_TABLES_EXT=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu ) > gpurun_out/t15_pytest_ext.log 2>&1; tail -n 3 gpurun_out/t15_pytest_ext.log
for v in 0 1 0 1; do BNK_TABLES_EXT=$v python bench.py --no-e2e --no-cpu --sub c4 --steps 10 > gpurun_out/t15_bench_ext$v.json 2>> gpurun_out/t15.err; python -c "
import json; d=json.load(open('gpurun_out/t15_bench_ext$v.json')); print('ext$v', d['ms_per_step'], d['phase_ms']['tables'], d['sub_results']['c4']['phase_ms']['tables'], d['parity']['joint_equal'], d['parity']['post_max_rel'])"; done
for c in 625 1250; do python bench.py --cols $c --no-e2e --no-cpu --no-sub --steps 10 > gpurun_out/t15_bench_cols$c.json 2>> gpurun_out/t15.err; python -c "
import json; d=json.load(open('gpurun_out/t15_bench_cols$c.json')); print('cols$c', d['ms_per_step'], d['phase_ms'], d['parity'])"; done
tail -c 300 gpurun_out/t15.err
