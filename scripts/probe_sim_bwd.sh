#!/bin/bash
# fused vs two-step phase 2 of the contrastive backward
for mode in fused twostep; do
  for d in 512 256; do
    echo "== MMDTI_SIM_BWD=$mode D=$d"
    MMDTI_SIM_BWD=$mode python scripts/bench_sim.py --nmin 2048 --nmax 32768 --d $d 2>&1 | grep -E "infonce|conr" | grep -E "N=  2048|N=  8192|N= 32768"
  done
done
