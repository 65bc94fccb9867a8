python scripts/probe_ew.py > /dev/null 2>&1 || exit 1
for ln in 1 2 3 4 6; do for rm in 2 4 8; do
  if [ $ln != 2 ] && [ $rm != 2 ]; then continue; fi
  MMDTI_LN_CTAS_PER_SM=$ln MMDTI_ROWMAP_CTAS_PER_SM=$rm ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ew_${ln}_${rm}.csv python scripts/probe_ew.py > /dev/null 2>&1
done; done
