mkdir -p gpurun_out/prof
timeout 300 python -m pytest tests/test_gpu_fds.py tests/test_gpu_edge_cases.py tests/test_gpu_hot_path_step.py -q 2>&1 | tail -3
python scripts/probe_fds.py > gpurun_out/prof/r2_fds_probe.log 2>&1; cat gpurun_out/prof/r2_fds_probe.log
python scripts/probe_fds.py --iters 2 > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:"fds_" -s 18 -c 9 -o gpurun_out/r2_fds python scripts/probe_fds.py --iters 2 > gpurun_out/ncu_fds.log 2>&1
python scripts/ncu_key_metrics.py gpurun_out/r2_fds.ncu-rep > gpurun_out/prof/r2_fds_ncu_full.txt 2>&1; rm -f gpurun_out/r2_fds.ncu-rep
grep -E "^==|time_duration|dram__bytes|dram_throughput" gpurun_out/prof/r2_fds_ncu_full.txt | cut -c1-120 | head -50
