#!/bin/bash
# usage: tools/ab_variants.sh name ... -- alternating bench runs of the current library and variants/lib_<name>.so
mkdir -p gpurun_out
for r in 1 2; do
  timeout 200 tools/ab.sh cur_$r
  for v in "$@"; do NBIOT_LIB=$PWD/variants/lib_$v.so timeout 200 tools/ab.sh ${v}_$r; done
done
