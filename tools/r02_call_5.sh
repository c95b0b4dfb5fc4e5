set -x
mkdir -p gpurun_out/r2e
out=gpurun_out/r2e/ab.txt; : > $out
V=$PWD/learn_path_tracing_b200/variants
for w in yoimiya_1080p 10_final_720p zhongli_4k intersect_10m; do
  tools/sweep.sh $w "" default >> $out
  for v in post rootbox postroot libm unit8 unit32; do
    PT_LIB_PATH=$V/libb200pt_$v.so tools/sweep.sh $w "" $v >> $out
  done
  tools/sweep.sh $w "" default >> $out
done
for w in yoimiya_1080p 10_final_720p; do
  for sm in 16 20 26; do tools/sweep.sh $w "--shade-min $sm" shade$sm >> $out; done
  for sv in 4 6 12; do tools/sweep.sh $w "--serve-min $sv" serve$sv >> $out; done
  for sm in 20 26; do PT_LIB_PATH=$V/libb200pt_post.so tools/sweep.sh $w "--shade-min $sm --serve-min 12" post_shade${sm}_serve12 >> $out; done
done
cat $out
# correctness of the variants that might become default: parity subset
for v in post rootbox postroot; do
  PT_LIB_PATH=$V/libb200pt_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_legacy.py -m gpu -x -q > gpurun_out/r2e/tests_$v.txt 2>&1; tail -2 gpurun_out/r2e/tests_$v.txt
done
