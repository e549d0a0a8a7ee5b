#!/bin/bash
# temporal-blocking kernel variants built into wayverb_b200/_exp (THREADS_TX_MINB), same box
for lib in wayverb_b200/_exp/libtb_*.so; do
  echo "== $lib"
  WVB_LIB=$PWD/$lib timeout 300 python tools/tb2_time.py 2>&1 | grep -E "tb2|identical|Error|error" | tail -3
done
