#!/bin/bash
# A/B of k_walk3's two modes on the C2 bench (device-resident part only), same box, one run per argument.
# usage: walk_ab.sh MODE ...   MODE = WR_WALK_PREFETCH: "" adaptive (default: wandering mode until the colony has converged),
#                              0 = predicted gathers always, 1 = neighbour-row prefetch + four-entry probes always
for mode in "$@"; do
  if [ -n "$mode" ]; then export WR_WALK_PREFETCH=$mode; else unset WR_WALK_PREFETCH; fi
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub --no-k26 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('mode=$mode value=%.4g ms/iter=%s steps/ant=%.0f e2e=%.4g'%(d['value'],{k:round(v,4) for k,v in d['kernel_ms_per_iteration'].items()},d['mean_steps_per_ant'],d['e2e']['value']))
"
done
