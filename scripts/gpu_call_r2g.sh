#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "splat or photon" 2>&1 | tail -2
python scripts/photon_sweep.py 4194304,67108864 2>/dev/null | cut -c1-330
EVPLP_LIB=$PWD/evplp_b200/lib/libevplp_b200_splat2.so python scripts/photon_sweep.py 4194304,67108864 2>/dev/null | cut -c1-330
scripts/quick_variants.sh "--no-single-frame" "--opt gather_cluster_size=8" "--opt gather_cluster_skip_max=16" "--opt gather_cluster_skip_max=256" "--opt gather_lpt=0"
