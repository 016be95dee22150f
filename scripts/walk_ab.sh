#!/bin/bash
# A/B of the walk kernel generations / prefetch modes / L2 warm-up on the C2 bench (device-resident part only).
# usage: walk_ab.sh "V PREFETCH WARM [CARVEOUT%]" ...
for cfg in "$@"; do
  set -- $cfg
  if [ -n "$4" ]; then export WR_WALK_CARVEOUT=$4; else unset WR_WALK_CARVEOUT; fi
  WR_WALK_V=$1 WR_WALK_PREFETCH=$2 WR_WALK_WARM=$3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('walk_v=$1 prefetch=$2 warm=$3 carveout=$4 value=%.4g ms/iter=%s steps/ant=%.0f e2e=%.4g'%(d['value'],{k:round(v,4) for k,v in d['kernel_ms_per_iteration'].items()},d['mean_steps_per_ant'],d['e2e']['value']))
"
done
