#!/bin/bash
# A/B of PT_MODE_PERSIST (3) against PT_MODE_DUAL (5) on the three render workloads -> gpurun_out/ab_dual.txt
out=gpurun_out/ab_dual.txt; : > $out
for w in 8_refract_1080p 10_final_720p yoimiya_1080p; do
  tools/sweep.sh $w "--mode 3" persist >> $out
  for k in 3 4; do for sm in 12 20 26; do
    tools/sweep.sh $w "--mode 5 --k $k --shade-min $sm" dual_b${k}_s${sm} >> $out
  done; done
done
cat $out
