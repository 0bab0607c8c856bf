#!/bin/bash
python -m pytest tests/test_gpu_cluster.py -m gpu -x -q 2>&1 | tail -2
scripts/quick_variants.sh "--no-single-frame" "" "--opt gather_cluster_extent_permille=40" "--opt gather_cluster_extent_permille=70" "--opt gather_cluster_extent_permille=120" "--opt gather_cluster_extent_permille=250"
