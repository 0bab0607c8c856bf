#!/bin/bash
# Round-2 ncu captures (one GPU).  usage: scripts/ncu_r2.sh [tag]
#  1. --set full of the kernel bench.py's headline runs (gather_cluster_kernel) on a launch short enough for ncu's ~40 replays
#     (full resolution, 1/16 of the VPL paths);
#  2. --set full of the BVH-build kernels, gather_vsl_kernel (config-3 scene), gather_lvc_kernel and path_trace_kernel.
tag=${1:-r2}
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:gather_cluster -c 1 -o gpurun_out/${tag}_cluster python bench.py --steps 1 --warmup 0 --no-cpu --no-single-frame --vpl-paths 1024 > gpurun_out/${tag}_cluster.log 2>&1
timeout 600 $NCU -k regex:'prim_bounds|morton|leaf_records|karras|refit|collapse|DeviceRadixSort' -c 40 -o gpurun_out/${tag}_bvh python scripts/config_runs.py C4 > gpurun_out/${tag}_bvh.log 2>&1
timeout 600 $NCU -k regex:gather_vsl -c 1 -o gpurun_out/${tag}_vsl python scripts/config_runs.py C3s > gpurun_out/${tag}_vsl.log 2>&1
timeout 600 $NCU -k regex:gather_lvc -c 1 -o gpurun_out/${tag}_lvc python scripts/config_runs.py LVC > gpurun_out/${tag}_lvc.log 2>&1
timeout 600 $NCU -k regex:path_trace -c 1 -o gpurun_out/${tag}_pt python scripts/config_runs.py PT > gpurun_out/${tag}_pt.log 2>&1
tail -n 3 gpurun_out/${tag}_*.log
