# dress rehearsal of the driver's round-end sequence on the final tree
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/t19_smoke.log 2>&1; tail -n 2 gpurun_out/t19_smoke.log
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/t19_pytest.log 2>&1; tail -n 6 gpurun_out/t19_pytest.log
( time python bench.py --impl reference ) > gpurun_out/t19_bench_ref.json 2> gpurun_out/t19_bench_ref.err; tail -n 4 gpurun_out/t19_bench_ref.err
( time python bench.py ) > gpurun_out/t19_bench.json 2> gpurun_out/t19_bench.err; tail -n 4 gpurun_out/t19_bench.err
python -c "
import json
r=json.load(open('gpurun_out/t19_bench_ref.json')); d=json.load(open('gpurun_out/t19_bench.json'))
print('ref', r['value'], r['ms_per_step'], r['cpu_baseline']['sample'][:80])
print('b200', d['value'], d['ms_per_step'], d['phase_ms']); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['per_pass']['marginal_all_ancestors']['ms'], d['e2e']['marginal_floor_ms_at_that_rate'])
print('ratio e2e', d['e2e']['value']/r['value'], 'device', d['value']/r['value']); print(d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['per_kernel']); print(d['clocks'], d['parity'], d['gpu_launches'])"
