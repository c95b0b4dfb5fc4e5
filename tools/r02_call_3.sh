set -x
mkdir -p gpurun_out/r2c
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2c/gpu_tests.txt 2>&1; tail -25 gpurun_out/r2c/gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c/smoke.txt 2>&1; tail -2 gpurun_out/r2c/smoke.txt
( time timeout 900 python bench.py > gpurun_out/r2c/bench_default.json 2> gpurun_out/r2c/bench_default.err ) 2>&1 | tail -3
tail -c 600 gpurun_out/r2c/bench_default.err
