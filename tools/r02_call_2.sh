set -x
mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b/gpu_tests.txt 2>&1; tail -3 gpurun_out/r2b/gpu_tests.txt
out=gpurun_out/r2b/ab.txt; : > $out
V=$PWD/learn_path_tracing_b200/variants
for w in yoimiya_1080p 10_final_720p 8_refract_1080p intersect_10m; do
  PT_NO_INLINE_TRIS=1 PT_LIB_PATH=$V/libb200pt_base.so tools/sweep.sh $w "" base >> $out
  tools/sweep.sh $w "" default >> $out
  PT_NO_INLINE_TRIS=1 tools/sweep.sh $w "" default_noinltri >> $out
  for v in ss8 ss24 pf glut; do
    PT_LIB_PATH=$V/libb200pt_$v.so tools/sweep.sh $w "" $v >> $out
  done
  PT_NO_INLINE_TRIS=1 PT_LIB_PATH=$V/libb200pt_base.so tools/sweep.sh $w "" base >> $out
done
cat $out
