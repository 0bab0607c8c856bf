#!/bin/bash
# quick GPU check: parity tests + short gather timings (profiling override, not the headline workload)
# usage: scripts/quick_gpu.sh "opt1=v,opt2=v" "opt=v" ...   (each argument = one bench run with those options)
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -2
for o in "$@"; do
  flags=""; IFS=',' read -ra parts <<< "$o"; for p in "${parts[@]}"; do flags="$flags --opt $p"; done
  python bench.py --steps 2 --warmup 1 --no-cpu --vpl-paths 512 $flags 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$o pairs/s %.4g rays/s %.4g' % (d['value'], d['shadow_rays_per_s']), d['stage_ms_per_step_rank0'], d.get('shaft_gather'))"
done
