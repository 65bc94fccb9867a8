python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
for f in none all in fc1 dfc2 "dout,din" wgrad out fc2 dfc1; do
  MMDTI_FUSED=$f python bench.py --workload encoder --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_$f.json 2> gpurun_out/b_$f.err; echo "fused=$f rc=$? $(python -c "import json;d=json.load(open('gpurun_out/b_$f.json'));print(d['ms_per_step'])")"
done
