#!/bin/bash
export WVB_LIB=$PWD/wayverb_b200/libwvb200_dbg.so
for cfg in "0 1" "1 1" "1 2"; do set -- $cfg
  echo "== B1SMALL=$1 blocks/SM=$2: $(WVB_WG_TB2_B1SMALL=$1 WVB_WG_TB2_B1BLOCKS=$2 timeout 300 python tools/tb2_time.py | grep -E "tb2|identical" | tail -2 | tr '\n' ' ')"
done
