set -x
mkdir -p gpurun_out/r2q
out=gpurun_out/r2q/ab.txt; : > $out
V=$PWD/learn_path_tracing_b200/variants
for w in yoimiya_1080p zhongli_4k; do
  tools/sweep.sh $w "" default_straight_taps >> $out
  PT_LIB_PATH=$V/libb200pt_brtaps.so tools/sweep.sh $w "" branchy_taps >> $out
  tools/sweep.sh $w "" default_straight_taps >> $out
  PT_LIB_PATH=$V/libb200pt_brtaps.so tools/sweep.sh $w "" branchy_taps >> $out
done
cat $out
timeout 600 python -m pytest tests/test_gpu_legacy.py tests/test_gpu_round2.py -m gpu -q > gpurun_out/r2q/tests.txt 2>&1; tail -n 3 gpurun_out/r2q/tests.txt
