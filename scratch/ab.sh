#!/bin/bash
# A/B of env toggles on the bench step: usage ab.sh "VAR=val ..." "VAR=val ..."
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-decode 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
b=d['breakdown']
print('ms/step %.3f' % d['ms_per_step'], ' '.join('%s=%.3f'%(k,v['ms_per_step']) for k,v in b.items()))
"
done
