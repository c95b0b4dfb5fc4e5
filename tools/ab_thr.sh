#!/bin/bash
out=gpurun_out/ab_thr.txt; : > $out
for w in yoimiya_1080p; do for sm in 18 22 26; do for sv in 6 8 12; do
  tools/sweep.sh $w "--shade-min $sm --serve-min $sv" shade${sm}_serve${sv} >> $out
done; done; done
for sm in 18 26; do tools/sweep.sh 10_final_720p "--shade-min $sm" shade${sm} >> $out; done
cat $out
