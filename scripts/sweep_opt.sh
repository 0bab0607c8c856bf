#!/bin/bash
# A/B of tuning options on a reduced gather (profiling override, not the headline workload)
# usage: scripts/sweep_opt.sh VPL_PATHS "opt1=v,opt2=v" "opt=v" ...
vp=$1; shift
for o in "$@"; do
  flags=""; IFS=',' read -ra parts <<< "$o"; for p in "${parts[@]}"; do [ -n "$p" ] && flags="$flags --opt $p"; done
  python bench.py --steps 2 --warmup 1 --no-cpu --vpl-paths $vp $flags 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$o] pairs/s %.4g gather_ms %.1f' % (d['value'], d['stage_ms_per_step_rank0']['vpl_gather']), d.get('shaft_gather'))"
done
