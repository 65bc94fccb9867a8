timeout 600 python -m pytest tests/test_gpu_cross_modal.py tests/test_gpu_chemberta.py tests/test_gpu_mm_model_dropin.py tests/test_cabi.py -x -q 2>&1 | grep -v Warning | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_hotpath_fusion_n1.json 2> gpurun_out/fus.err; tail -3 gpurun_out/fus.err
python bench.py --chemberta --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_hotpath_chemberta_n1.json 2> gpurun_out/cb.err; tail -3 gpurun_out/cb.err
for f in hotpath_fusion_n1 hotpath_chemberta_n1; do python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['gpu_launches'])"; done
