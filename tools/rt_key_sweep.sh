#!/bin/bash
# sort-key sweep of the wavefront ray loop (debug-knob library)
export WVB_LIB=$PWD/wayverb_b200/libwvb200_dbg.so
for cfg in "5 3" "5 2" "4 4" "4 3" "5 4" "3 5" "4 5"; do
  set -- $cfg
  for sc in hall box; do
    echo "side_bits=$1 dir_bits=$2 $sc: $(WVB_RT_KEY_SIDE_BITS=$1 WVB_RT_KEY_DIR_BITS=$2 timeout 200 python tools/profile_rt.py 1048576 $sc | grep wavefront | tail -1)"
  done
done
