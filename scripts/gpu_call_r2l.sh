#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_baseline_configs.py tests/test_gpu_reference.py tests/test_gpu_host_cli.py -m gpu -x -q 2>&1 | tail -3
python scripts/photon_sweep.py 4194304,67108864 2>/dev/null | cut -c1-330
scripts/quick_variants.sh "--no-single-frame" ""
python -c "
import json; d=json.load(open('gpurun_out/q.json')); print(d['stage_ms_per_step_rank0'])"
