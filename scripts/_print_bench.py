import json,sys
d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["stage_ms_per_step_rank0"]["vpl_gather"])
