set -x
mkdir -p gpurun_out/r2l
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29547"
( time timeout 900 $TR bench.py --gpus 8 > gpurun_out/r2l/n8_default.json 2> gpurun_out/r2l/n8_default.err ) 2>&1 | tail -n 3
timeout 300 $TR bench.py --gpus 8 --workload yoimiya_1080p --only --steps 5 > gpurun_out/r2l/n8_yoimiya.json 2> gpurun_out/r2l/n8_yoimiya.err
timeout 200 $TR bench.py --gpus 8 --impl reference --steps 1 --warmup 0 --workload 9_dof_720p > gpurun_out/r2l/n8_reference.json 2> gpurun_out/r2l/n8_reference.err
tail -c 300 gpurun_out/r2l/n8_default.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2l/n8_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), d['unit'], round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d.get('n_gpus'))
        for n,e in d.get('workloads',{}).items(): print('   ',n, round(e['value']), e['unit'], round(e['ms_per_step'],3), 'e2e', round(e['e2e']['value']))
    except Exception as ex: print(f,'ERR',ex)
PY
