N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { tag=$1; shift; env "$@" timeout 200 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/t_$tag.json 2> gpurun_out/t_$tag.err; python -c "
import json
d=json.loads(open('gpurun_out/t_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value']), round(d['ms_per_step'],3))"; }
run base A=1
run b16 MMDTI_BUCKET_MB=16
run b64 MMDTI_BUCKET_MB=64
run cta4 NCCL_MAX_CTAS=4
run cta8 NCCL_MAX_CTAS=8
run cta16 NCCL_MAX_CTAS=16
run fp32 MMDTI_GRAD_COMM=fp32
