#!/bin/bash
# ncu --set full of the K2 kernels at the large-molecule shape (L = 258, B = 32), both forms of the forward.
# Run under gpurun on ONE GPU after `python scripts/probe_k2.py --L 258 --B 32` has exited 0 without ncu:
#   gpurun --timeout 400 -- 'bash scripts/ncu_k2_large_l.sh'
# then here:  python scripts/ncu_key_metrics.py gpurun_out/k2_l258_rowsplit.ncu-rep ; python scripts/ncu_lines.py <rep> pair_attn_fwd 30
set -e
mkdir -p gpurun_out
python scripts/probe_k2.py --L 258 --B 32 --iters 3 > gpurun_out/probe_k2_l258.log 2>&1
for cs in 0 1; do
  name=$([ "$cs" = 1 ] && echo colsplit || echo rowsplit)
  MMDTI_K2_FWD_CS=$cs timeout 180 ncu --set full --clock-control none --import-source on -k regex:pair_attn -s 6 -c 2 \
      -o gpurun_out/k2_l258_$name python scripts/probe_k2.py --L 258 --B 32 --iters 3 > gpurun_out/ncu_k2_l258_$name.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
