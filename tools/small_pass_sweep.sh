#!/bin/bash
# k_raster's thread-per-record pass (RXC_SMALL_GSHIFT / RXC_SMALL_MIN_LIST / RXC_SMALL_MAX_PIX): device-resident times of the
# dense 8K frame for a few settings; min_list 0 = pass off.
for cfg in ${SWEEP:-"3 0 64" "0 128 64" "2 128 64" "2 64 64" "2 64 128" "3 128 64" "3 64 64" "3 32 64" "3 64 128" "3 64 256" "3 32 256" "4 64 128" "4 64 256" "4 32 512"}; do
  set -- $cfg
  echo -n "gshift $1 min_list $2 max_pix $3: "
  RXC_SMALL_GSHIFT=$1 RXC_SMALL_MIN_LIST=$2 RXC_SMALL_MAX_PIX=$3 python tools/quick_bench.py dense8k 2>&1 | tail -1
done
