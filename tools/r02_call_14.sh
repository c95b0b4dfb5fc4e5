set -x
mkdir -p gpurun_out/r2n
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2n/gpu_tests.txt 2>&1; tail -n 4 gpurun_out/r2n/gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n/smoke.txt 2>&1; tail -n 2 gpurun_out/r2n/smoke.txt
bash tools/ncu_capture.sh > gpurun_out/r2n/ncu_capture.log 2>&1; tail -n 6 gpurun_out/r2n/ncu_capture.log
