set -x
mkdir -p gpurun_out/r2f
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multigpu.py tests/test_multigpu_gloo.py -m "gpu or not gpu" -q -s > gpurun_out/r2f/multigpu_tests.txt 2>&1; tail -6 gpurun_out/r2f/multigpu_tests.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for b in 1 4 8; do
  timeout 300 $TR bench.py --gpus 2 --workload 8_refract_1080p --only --steps 5 --bands $b > gpurun_out/r2f/n2_refract_bands$b.json 2> gpurun_out/r2f/n2_refract_bands$b.err
  timeout 300 $TR bench.py --gpus 2 --workload yoimiya_1080p --only --steps 3 --bands $b > gpurun_out/r2f/n2_yoimiya_bands$b.json 2> gpurun_out/r2f/n2_yoimiya_bands$b.err
done
( time timeout 900 $TR bench.py --gpus 2 > gpurun_out/r2f/n2_default.json 2> gpurun_out/r2f/n2_default.err ) 2>&1 | tail -3
timeout 300 $TR bench.py --gpus 2 --impl reference --steps 1 --warmup 0 --workload 8_refract_1080p > gpurun_out/r2f/n2_reference.json 2> gpurun_out/r2f/n2_reference.err
tail -c 400 gpurun_out/r2f/n2_default.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f/n2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), d['unit'], round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
        for n,e in d.get('workloads',{}).items(): print('   ',n, round(e['value']), e['unit'], round(e['ms_per_step'],3), 'e2e', round(e['e2e']['value']))
    except Exception as ex: print(f,'ERR',ex)
PY
