#!/bin/bash
# usage: tools/sweep.sh <workload> "<extra args>" label   -> prints one line
w=$1; shift; extra=$1; shift; label=$1
timeout 300 python bench.py --workload $w --only --steps 3 --no-cpu $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w', '$label', round(d['value']), d['unit'], round(d['ms_per_step'],2), 'ms', 'mrays', round(d.get('mrays_per_s',0)))"
