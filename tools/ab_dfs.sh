#!/bin/bash
# A/B of the depth-first renumbering of PLOC trees (PT_PLOC_DFS=0 keeps creation order) -> gpurun_out/ab_dfs.txt
out=gpurun_out/ab_dfs.txt; : > $out
for w in "$@"; do for d in 0 1 0 1; do
  PT_PLOC_DFS=$d timeout 300 python bench.py --workload $w --only --steps 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d.get('roofline_fp32') or {}
print('$w', 'dfs=$d', round(d['value']), d['unit'], round(d['ms_per_step'],2), 'ms', 'e2e', round(d['e2e']['value']), 'nodes/seg', round(f.get('nodes_per_segment',0),2))" >> $out
done; done
cat $out
