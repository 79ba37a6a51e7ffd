#!/bin/bash
# k_raster's thread-per-record pass (RXC_SMALL_MIN_LIST / RXC_SMALL_MAX_PIX): device-resident times of the dense 8K frame
# and of the map scene for a few settings; 0 = pass off.
for cfg in "0 64" "32 64" "64 64" "128 64" "256 64" "128 16" "128 32" "128 128" "64 256" "128 1024"; do
  set -- $cfg
  echo "== min_list $1 max_pix $2"
  RXC_SMALL_MIN_LIST=$1 RXC_SMALL_MAX_PIX=$2 python tools/quick_bench.py dense8k teapot1080 2>&1 | tail -2
done
