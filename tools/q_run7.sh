#!/bin/bash
# is the warp-per-instance kernel of this tree as fast as the committed one? (same box, alternating)
mkdir -p gpurun_out
for r in 1 2; do
  NBIOT_LIB=$PWD/variants/lib_head.so timeout 200 tools/ab.sh head_$r
  timeout 200 tools/ab.sh cur_$r
done
