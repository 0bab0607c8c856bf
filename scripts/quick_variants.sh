#!/bin/bash
# quick A/B of gather variants on a reduced workload (same VPL density as the headline, 1/4 of the pixels)
# usage: quick_variants.sh "<bench args>" "opts1" "opts2" ...
base=$1; shift
for o in "$@"; do
  timeout 90 python bench.py --steps 2 --warmup 2 --no-cpu $base $o > gpurun_out/q.json 2> gpurun_out/q.err
  python - "$o" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/q.json"))
    sg, cg = d.get("shaft_gather") or {}, d.get("cluster_gather") or {}
    print("%-40s pairs/s %.4g gather_ms %.1f | nodes/step %.2f cand/step %.2f packets %.4f | desc/step %s cand/desc %s" % (
        sys.argv[1], d["value"], d["stage_ms_per_step_rank0"]["vpl_gather"], sg.get("nodes_per_step", 0), sg.get("candidate_leaves_per_step", 0),
        sg.get("fallback_frac", 0), cg.get("descents_per_step"), cg.get("candidates_per_descent")))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open("gpurun_out/q.err").read()[-600:])
PY
done
