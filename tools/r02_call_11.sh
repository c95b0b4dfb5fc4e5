set -x
mkdir -p gpurun_out/r2k
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2k/gpu_tests.txt 2>&1; tail -n 14 gpurun_out/r2k/gpu_tests.txt
PT_LIB_PATH=$PWD/learn_path_tracing_b200/libb200pt_exp.so timeout 900 python -m pytest tests/test_gpu_experimental.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2k/gpu_tests_experimental_lib.txt 2>&1; tail -n 3 gpurun_out/r2k/gpu_tests_experimental_lib.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k/smoke.txt 2>&1; tail -n 2 gpurun_out/r2k/smoke.txt
( time timeout 900 python bench.py > gpurun_out/r2k/bench_default.json 2> gpurun_out/r2k/bench_default.err ) 2>&1 | tail -n 3
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2k/bench_reference.json 2> gpurun_out/r2k/bench_reference.err ) 2>&1 | tail -n 3
bash tools/ncu_capture.sh > gpurun_out/r2k/ncu_capture.log 2>&1; tail -n 8 gpurun_out/r2k/ncu_capture.log
