mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/t21_bench_n4.json 2> gpurun_out/t21_bench_n4.err; tail -c 400 gpurun_out/t21_bench_n4.err
python -c "
import json; d=json.load(open('gpurun_out/t21_bench_n4.json')); print('n4', d['ms_per_step'], d['value'], d['phase_ms']); print(json.dumps(d['e2e'])[:500]); print({k:(v.get('ms_per_step'), v.get('ms_per_evaluation'), v.get('joint_plus_state_gather_ms')) for k,v in d['sub_results'].items()})"
