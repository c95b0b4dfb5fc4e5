# record of GPU call 1 of round 2 (ran against the round-1 bench.py, which still had --wide; the 4-wide walk now lives in libb200pt_exp.so)
set -x
mkdir -p gpurun_out/r2a
PT_TEST_WIDE_RENDER=1 timeout 600 python -m pytest tests -m gpu -x -q -k "wide" > gpurun_out/r2a/wide_tests.txt 2>&1
for w in yoimiya_1080p 10_final_720p zhongli_4k; do
  timeout 300 python bench.py --workload $w --steps 3 --no-cpu > gpurun_out/r2a/${w}_base.json 2> gpurun_out/r2a/${w}_base.err
  timeout 300 python bench.py --workload $w --steps 3 --no-cpu --wide > gpurun_out/r2a/${w}_wide.json 2> gpurun_out/r2a/${w}_wide.err
done
timeout 400 python bench.py --workload intersect_10m --steps 3 > gpurun_out/r2a/int_base.json 2> gpurun_out/r2a/int_base.err
timeout 400 python bench.py --workload intersect_10m --steps 3 --wide > gpurun_out/r2a/int_wide.json 2> gpurun_out/r2a/int_wide.err
tail -3 gpurun_out/r2a/wide_tests.txt
