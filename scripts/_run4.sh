timeout 600 python -m pytest tests/test_gpu_chemberta.py tests/test_gpu_mm_model_dropin.py -x -q 2>&1 | grep -v Warning | tail -8
grep "chemberta\|mm_model" gpurun_out/test_report.txt | cut -c1-1200
python bench.py --chemberta --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_hotpath_chemberta_n1.json 2> gpurun_out/cb.err; tail -3 gpurun_out/cb.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_hotpath_chemberta_n1.json').read().strip().splitlines()[-1]); print('chemberta', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['gpu_launches'], d['config']['workload'])"
