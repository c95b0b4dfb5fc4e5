set -x
mkdir -p gpurun_out/r2o
( time timeout 900 python bench.py > gpurun_out/r2o/bench_default.json 2> gpurun_out/r2o/bench_default.err ) 2>&1 | tail -n 3
( time timeout 600 python bench.py --impl reference > gpurun_out/r2o/bench_reference.json 2> gpurun_out/r2o/bench_reference.err ) 2>&1 | tail -n 3
tail -c 300 gpurun_out/r2o/bench_default.err
