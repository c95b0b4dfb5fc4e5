set -x
mkdir -p gpurun_out/r2y
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561"
( time timeout 600 $TR bench.py --gpus 4 > gpurun_out/r2y/n4_default.json 2> gpurun_out/r2y/n4_default.err ) 2>&1 | tail -n 3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y/n4_default.json').read().strip().splitlines()[-1]); print(d['config']['workload'], round(d['value']), d['unit'], round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['n_gpus'])
for n,e in d.get('workloads',{}).items(): print('   ',n, round(e['value']), e['unit'], round(e['ms_per_step'],3), 'e2e', round(e['e2e']['value']))
PY
