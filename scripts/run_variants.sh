#!/bin/bash
# usage: run_variants.sh tag "opts1" "opts2" ...   (each opts = space-separated --opt k=v)
tag=$1; shift
i=0
for o in "$@"; do
  timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu $o > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_$i.json"))
    print("$o", "| pairs/s %.4g"%d["value"], "| gather ms %.1f"%d["stage_ms_per_step_rank0"]["vpl_gather"], "|", d.get("shaft_gather"), d.get("cluster_gather"))
except Exception as e:
    print("$o FAILED", e); print(open("gpurun_out/${tag}_$i.err").read()[-800:])
PY
  i=$((i+1))
done
