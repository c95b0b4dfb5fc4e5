set -x
mkdir -p gpurun_out/r2j
out=gpurun_out/r2j/ab.txt; : > $out
V=$PWD/learn_path_tracing_b200/variants
for w in yoimiya_1080p 10_final_720p zhongli_4k; do
  tools/sweep.sh $w "" default >> $out
  for v in rs16 rs8; do PT_LIB_PATH=$V/libb200pt_$v.so tools/sweep.sh $w "" $v >> $out; done
  tools/sweep.sh $w "" default >> $out
done
cat $out
