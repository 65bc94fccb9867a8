#!/bin/bash
# phase-1 / phase-2 sensitivity of the tcgen05 similarity kernel to the number of ring stages and to D
for st in 2 3 5; do
  echo "== stages=$st D=512"; MMDTI_SIM_STAGES=$st python scripts/bench_sim.py --nmin 32768 --nmax 32768 --d 512 2>&1 | grep -E "infonce|conr"
done
for d in 128 256; do
  echo "== stages=max D=$d"; python scripts/bench_sim.py --nmin 32768 --nmax 32768 --d $d 2>&1 | grep -E "infonce|conr"
done
