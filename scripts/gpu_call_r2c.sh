#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_cluster.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
scripts/quick_variants.sh "--no-single-frame" "" "--opt bvh_leaf_max=1" "--opt gather_min_blocks=2" "--opt gather_min_blocks=4"
python scripts/config_runs.py C3s 2>&1 | tail -1 | cut -c1-300
EVPLP_LIB=$PWD/evplp_b200/lib/libevplp_b200_ni.so python scripts/config_runs.py C3s 2>&1 | tail -1 | cut -c1-300
