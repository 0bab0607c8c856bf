#!/bin/bash
python -m pytest tests/test_gpu_cluster.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
scripts/quick_variants.sh "--no-single-frame" "" "--opt gather_min_blocks=3"
