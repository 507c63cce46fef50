#!/bin/bash
# L2 prefetch distance sweep of the aggregation kernels (tuning aid): bash tools/agg_pf_sweep.sh
cd "$(dirname "$0")"
for cfg in "0 0" "32 6" "64 6" "96 6" "160 6" "32 3" "32 12" "32 24" "64 12"; do
  set -- $cfg
  echo "pfPixels=$1 pfRows=$2: $(CARTB200_PF_PIXELS=$1 CARTB200_PF_ROWS=$2 python agg_waves.py 2>&1 | tail -1)"
done
