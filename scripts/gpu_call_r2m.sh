#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_baseline_configs.py tests/test_gpu_reference.py -m gpu -x -q 2>&1 | tail -3
python scripts/photon_sweep.py 4194304,67108864 2>/dev/null > gpurun_out/sweep_r2_final_1gpu.jsonl; cut -c1-120,330-470 gpurun_out/sweep_r2_final_1gpu.jsonl
timeout 90 python bench.py --steps 1 --warmup 1 --no-cpu --no-single-frame > gpurun_out/q.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/q.json')); print(d['value'], d['stage_ms_per_step_rank0'])"
