#!/bin/bash
# aggregation kernel variant sweep (tuning aid): bash tools/agg_variant_sweep.sh
cd "$(dirname "$0")"
for cfg in "0 0" "6 0" "3 0" "4 0" "7 0" "8 0" "6 1" "6 2" "0 0"; do
  set -- $cfg
  echo "horiz=$1 vert=$2: $(CARTB200_HORIZ_VARIANT=$1 CARTB200_VERT_VARIANT=$2 python agg_waves.py 2>&1 | tail -1)"
done
