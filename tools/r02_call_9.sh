set -x
mkdir -p gpurun_out/r2i
out=gpurun_out/r2i/ab.txt; : > $out
V=$PWD/learn_path_tracing_b200/variants
for w in yoimiya_1080p 10_final_720p zhongli_4k intersect_10m; do
  tools/sweep.sh $w "" default >> $out
  for v in steps3 specpop; do PT_LIB_PATH=$V/libb200pt_$v.so tools/sweep.sh $w "" $v >> $out; done
done
for sm in 16 18 22; do for sv in 6 8 10; do tools/sweep.sh yoimiya_1080p "--shade-min $sm --serve-min $sv" shade${sm}_serve$sv >> $out; done; done
for sm in 20 24; do for sv in 10 14 16; do tools/sweep.sh 10_final_720p "--shade-min $sm --serve-min $sv" shade${sm}_serve$sv >> $out; done; done
# batch kernel: serve_min (bits 8-13) x fetch_min (bits 14-19)
for sv in 4 8 12 16; do for fm in 8 16 24; do tools/sweep.sh intersect_10m "--trace-flags $(( (sv<<8) | (fm<<14) ))" serve${sv}_fetch$fm >> $out; done; done
cat $out
