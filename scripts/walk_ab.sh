#!/bin/bash
# A/B of the pass-1 walk kernels on the C2 bench (device-resident part only), same box, alternating.
# usage: walk_ab.sh "V [PREFETCH]" ...      V = WR_WALK_V (2 = k_walk2, 3 = k_walk3), PREFETCH = WR_WALK_PREFETCH (default: adaptive)
for cfg in "$@"; do
  set -- $cfg
  if [ -n "$2" ]; then export WR_WALK_PREFETCH=$2; else unset WR_WALK_PREFETCH; fi
  WR_WALK_V=$1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub --no-k26 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('walk_v=$1 prefetch=$2 value=%.4g ms/iter=%s steps/ant=%.0f e2e=%.4g full_search=%.4g s'%(d['value'],{k:round(v,4) for k,v in d['kernel_ms_per_iteration'].items()},d['mean_steps_per_ant'],d['e2e']['value'],d.get('full_search',{}).get('seconds_from_host_buffers',0)))
"
done
