N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_hotpath_fusion_n$N.json 2> gpurun_out/nNa.err; tail -2 gpurun_out/nNa.err
timeout 300 $TR bench.py --gpus $N --workload config3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_config3_fusion_n$N.json 2> gpurun_out/nNb.err; tail -2 gpurun_out/nNb.err
for f in hotpath_fusion_n$N config3_fusion_n$N; do python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['gpu_launches'], d['n_gpus'], d['clocks'])"; done
