timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_elementwise.py tests/test_gpu_encoder.py tests/test_gpu_optim.py -x -q 2>&1 | tail -3
timeout 200 python scripts/probe_gemm.py 2>&1 | tail -14
python bench.py --workload encoder --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_all.json 2> gpurun_out/b_all.err; echo "encoder rc=$? $(python -c "import json;d=json.load(open('gpurun_out/b_all.json'));print(d['ms_per_step'])")"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_hot.json 2> gpurun_out/b_hot.err; echo "hotpath rc=$? $(python -c "import json;d=json.load(open('gpurun_out/b_hot.json'));print(d['ms_per_step'])")"
timeout 300 python bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b_hot2.json 2> gpurun_out/b_hot2.err; echo "n2 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/b_hot2.json'));print(d['ms_per_step'], d['value'])")"
