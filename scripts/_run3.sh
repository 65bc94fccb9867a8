set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -8
python bench.py --workload config4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_config4_fusion_n1.json 2> gpurun_out/c4.err; tail -2 gpurun_out/c4.err
python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_config3_fusion_n1.json 2> gpurun_out/c3.err; tail -2 gpurun_out/c3.err
python bench.py --workload encoder --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_encoder_n1b.json 2> gpurun_out/enc.err; tail -2 gpurun_out/enc.err
for f in config4_fusion_n1 config3_fusion_n1 encoder_n1b; do python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['gpu_launches'], d['roofline']['frac'])"; done
