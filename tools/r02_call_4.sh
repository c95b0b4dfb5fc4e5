set -x
mkdir -p gpurun_out/r2d
timeout 1800 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r2d/gpu_tests.txt 2>&1; tail -30 gpurun_out/r2d/gpu_tests.txt
bash tools/ncu_capture.sh > gpurun_out/r2d/ncu_capture.log 2>&1; tail -12 gpurun_out/r2d/ncu_capture.log
