mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu ) > gpurun_out/t14_pytest_multi.log 2>&1; tail -n 4 gpurun_out/t14_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/t14_bench_n2.json 2> gpurun_out/t14_bench_n2.err; tail -c 400 gpurun_out/t14_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/t14_bench_n2.json')); print('n2', d['ms_per_step'], d['value'], d['phase_ms']); print(json.dumps(d['e2e'])[:700]); print({k:(v.get('ms_per_step'), v.get('ms_per_evaluation')) for k,v in d['sub_results'].items()})"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/t14_bench_ref_n2.json 2> gpurun_out/t14_bench_ref_n2.err; tail -c 300 gpurun_out/t14_bench_ref_n2.json
