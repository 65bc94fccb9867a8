python scripts/probe_cross.py 2>&1 | tail -1
python scripts/probe_cross.py --B 32 --L1 258 --L2 256 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:cross_attn --launch-skip 18 --launch-count 6 -o gpurun_out/cross_r2 python scripts/probe_cross.py --iters 2 > gpurun_out/ncu_cross.log 2>&1
python scripts/ncu_key_metrics.py gpurun_out/cross_r2.ncu-rep > gpurun_out/r2_cross_attn_ncu_full.txt 2>&1
python scripts/ncu_lines.py gpurun_out/cross_r2.ncu-rep cross_attn 30 > gpurun_out/r2_cross_attn_ncu_source_top30.txt 2>&1
rm -f gpurun_out/cross_r2.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_launches_bench_hotpath_fusion_nograph.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
python scripts/ncu_launch_summary.py gpurun_out/r2_launches_bench_hotpath_fusion_nograph.csv > gpurun_out/r2_launches_bench_hotpath_fusion_nograph_summary.txt 2>&1
head -40 gpurun_out/r2_launches_bench_hotpath_fusion_nograph_summary.txt
grep -A8 "cross_attn" gpurun_out/r2_cross_attn_ncu_full.txt | head -80
