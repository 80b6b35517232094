#!/bin/bash
# fused two-phase experiment: timing at several lead distances + DRAM traffic / L2 hit rate of one launch
set -x
mkdir -p gpurun_out
for lead in ${LEADS:-320}; do
  CHIMP_TP_LEAD=$lead CHIMP_TP_FUSED=1 timeout 300 python scripts/measure_configs.py twophase 2>&1 | tail -1 | cut -c1-330
done
CHIMP_TP_FUSED=0 timeout 300 python scripts/measure_configs.py twophase 2>&1 | tail -1 | cut -c1-330
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:twoPhaseFused -s 3 -c 1 --csv --log-file gpurun_out/tp_fused_metrics.csv python scripts/measure_configs.py twophase > /dev/null 2>&1; cat gpurun_out/tp_fused_metrics.csv | tail -6 | cut -d, -f 5,11-20
