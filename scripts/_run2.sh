timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_elementwise.py tests/test_gpu_encoder.py -x -q 2>&1 | tail -3
timeout 200 python scripts/probe_gemm.py --only "fc" 2>&1 | tail -8
python bench.py --workload encoder --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_all.json 2> gpurun_out/b_all.err; echo "encoder rc=$? $(python -c "import json;d=json.load(open('gpurun_out/b_all.json'));print(d['ms_per_step'])")"
