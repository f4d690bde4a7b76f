#!/bin/bash
# usage: tools/dbg/sweep.sh "ENV1=a ENV2=b" "ENV1=c" ...   (one bench run per argument)
for cfg in "$@"; do
  ( for kv in $cfg; do export "$kv"; done
    python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys
t=sys.stdin.read().strip().splitlines()
if not t: print('$cfg', 'FAILED'); sys.exit()
d=json.loads(t[-1])
print('$cfg', 'value', round(d['value'],4), 'ms', round(d['ms_per_step']), {k:round(v) for k,v in d['stage_ms_per_step'].items() if v>1}, 'aligned', d['last_step_work']['pairs_aligned'], 'used', d['last_step_work']['alignments_consumed'])
" )
done
