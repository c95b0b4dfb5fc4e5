#!/bin/bash
out=gpurun_out/ab_radius.txt; : > $out
for w in yoimiya_1080p zhongli_4k; do for r in 4 8 16 32 64 128; do
  PT_PLOC_RADIUS=$r PT_BUILD_VERBOSE=1 timeout 300 python bench.py --workload $w --only --steps 2 --no-cpu 2>gpurun_out/ab_radius.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d.get('roofline_fp32') or {}
print('$w', 'radius $r', round(d['value']), d['unit'], round(d['ms_per_step'],2), 'ms', 'e2e', round(d['e2e']['value']), 'nodes/seg', round(f.get('nodes_per_segment',0),2))" >> $out
  grep -m1 "SAH cost" gpurun_out/ab_radius.err >> $out
done; done
cat $out
