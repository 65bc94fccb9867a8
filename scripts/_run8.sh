TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_hotpath_fusion_n8.json 2> gpurun_out/n8a.err; tail -2 gpurun_out/n8a.err
timeout 300 $TR bench.py --gpus 8 --workload config3 --steps 10 --warmup 3 > gpurun_out/r2_bench_config3_fusion_n8.json 2> gpurun_out/n8b.err; tail -2 gpurun_out/n8b.err
timeout 300 $TR bench.py --gpus 8 --workload config4 --steps 10 --warmup 3 > gpurun_out/r2_bench_config4_fusion_n8.json 2> gpurun_out/n8c.err; tail -2 gpurun_out/n8c.err
for f in hotpath_fusion_n8 config3_fusion_n8 config4_fusion_n8; do python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['gpu_launches'], d['n_gpus'], d['clocks'])"; done
