set -x
mkdir -p gpurun_out/r2g
out=gpurun_out/r2g/ab.txt; : > $out
V=$PWD/learn_path_tracing_b200/variants
for w in yoimiya_1080p 10_final_720p zhongli_4k intersect_10m; do
  tools/sweep.sh $w "" default >> $out
  for v in brless brless_t0; do
    PT_LIB_PATH=$V/libb200pt_$v.so tools/sweep.sh $w "" $v >> $out
  done
  tools/sweep.sh $w "" default >> $out
done
cat $out
PT_LIB_PATH=$V/libb200pt_brless.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_legacy.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r2g/tests_brless.txt 2>&1; tail -n 3 gpurun_out/r2g/tests_brless.txt
