#!/bin/bash
# A/B of the library variants built by tools/build_variants.sh -> gpurun_out/ab_variants.txt
out=gpurun_out/ab_variants.txt; : > $out
V=learn_path_tracing_b200/variants
for w in 8_refract_1080p 10_final_720p yoimiya_1080p; do
  for v in "$@"; do
    PT_LIB_PATH=$PWD/$V/libb200pt_$v.so tools/sweep.sh $w "" $v >> $out
  done
done
for v in base tos all; do
  [ -f $V/libb200pt_$v.so ] && PT_LIB_PATH=$PWD/$V/libb200pt_$v.so tools/sweep.sh intersect_10m "" $v >> $out
done
cat $out
