#!/bin/bash
# usage: tools/run_n.sh N workload [extra bench args]  -> gpurun_out/scale_<workload>_n<N>.json
N=$1; shift; W=$1; shift
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --workload $W --only "$@" > gpurun_out/scale_${W}_n$N.json 2> gpurun_out/scale_${W}_n$N.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --workload $W --only "$@" > gpurun_out/scale_${W}_n$N.json 2> gpurun_out/scale_${W}_n$N.err
fi
echo "$W n=$N rc=$?"
