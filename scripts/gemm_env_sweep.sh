#!/bin/bash
# library-GEMM settings (cuBLAS vs cuBLASLt dispatch, workspace) on the config-2 step
run() {
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g_$1.json 2> gpurun_out/g_$1.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/g_$1.json')); print('$1', round(d['value']), round(d['ms_per_step'],3))
except Exception as e:
    print('$1 ERR', e)"
}
run default
TORCH_BLAS_PREFER_CUBLASLT=1 run preferlt
CUBLASLT_WORKSPACE_SIZE=131072 CUBLAS_WORKSPACE_CONFIG=:131072:2 run bigws
TORCH_BLAS_PREFER_CUBLASLT=1 CUBLASLT_WORKSPACE_SIZE=131072 run preferlt_bigws
