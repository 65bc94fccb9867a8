#!/bin/bash
# 8-GPU gradient-exchange settings, one bench run each
run() {
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n8_$1.json 2> gpurun_out/n8_$1.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/n8_$1.json")); print("$1", round(d["value"]), round(d["ms_per_step"], 3))
except Exception as e:
    print("$1 ERR", e); print(open("gpurun_out/n8_$1.err").read()[-800:])
PY
}
run default
NCCL_MAX_CTAS=8 run ctas8
MMDTI_BUCKET_MB=96 run bucket96
