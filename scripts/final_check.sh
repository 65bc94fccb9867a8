# round-end check on one B200: GPU parity suite, smoke, the default bench line (with the reference arm beside it)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v Warning | tail -3
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/final.err; tail -2 gpurun_out/final.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1]); print('default', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['gpu_launches'], d['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['clocks'])"
