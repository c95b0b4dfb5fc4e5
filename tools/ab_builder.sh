#!/bin/bash
# A/B of the two hierarchy builders (PT_BUILDER=lbvh|ploc) -> gpurun_out/ab_builder.txt
out=gpurun_out/ab_builder.txt; : > $out
for w in "$@"; do for b in lbvh ploc ""; do
  PT_BUILD_VERBOSE=1 PT_BUILDER=$b timeout 300 python bench.py --workload $w --only --steps 3 --no-cpu 2>gpurun_out/ab_builder.err | tee /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d.get('roofline_fp32') or {}
print('$w', '$b' or 'auto', round(d['value']), d['unit'], round(d['ms_per_step'],2), 'ms', 'e2e', round(d['e2e']['value']), 'nodes/seg', round(f.get('nodes_per_segment',0),2), 'prims/seg', round(f.get('prims_per_segment',0),2))" >> $out
  grep -m1 "SAH cost" gpurun_out/ab_builder.err >> $out
done; done
cat $out
