#!/bin/bash
# quick GPU check: parity tests + short gather timings (profiling override, not the headline workload)
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for o in "$@"; do
python bench.py --steps 2 --warmup 1 --no-cpu --vpl-paths 512 --opt $o 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$o pairs/s %.4g rays/s %.4g' % (d['value'], d['shadow_rays_per_s']), d['stage_ms_per_step_rank0'])"
done
