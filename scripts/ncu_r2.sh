#!/bin/bash
# Round-2 ncu captures (one GPU).  usage: scripts/ncu_r2.sh [tag] [what...]
#  cluster : --set full of the kernel bench.py's headline runs (gather_cluster_kernel) on every 16th 8x4-pixel tile of the HEADLINE
#            frame with the full VPL set (the whole 3.4 s launch does not survive ncu's instrumented replays: two attempts timed out)
#  bvh / vsl / lvc / pt : --set full of the BVH-build kernels, gather_vsl_kernel (config-3 scene), gather_lvc_kernel, path_trace_kernel
# Only text summaries (scripts/ncu_summary.py, ncu_hot.py, ncu_regions.py) and the cluster report come back: gpurun_out/ is capped at 64 MiB.
tag=${1:-r2}; shift
what=${@:-cluster bvh vsl lvc pt}
NCU="ncu --set full --clock-control none --import-source on -f"
summ() {  # report, kernel regex
  python scripts/ncu_summary.py gpurun_out/$1.ncu-rep "$2" > gpurun_out/$1_summary.txt 2>&1
  python scripts/ncu_hot.py gpurun_out/$1.ncu-rep >> gpurun_out/$1_summary.txt 2>&1
  python scripts/ncu_regions.py gpurun_out/$1.ncu-rep 2.0 >> gpurun_out/$1_summary.txt 2>&1
}
for w in $what; do
  case $w in
    cluster) timeout 420 $NCU -k regex:gather_cluster -c 1 -o gpurun_out/${tag}_cluster python bench.py --steps 1 --warmup 0 --no-cpu --no-single-frame --tile-share 16 > gpurun_out/${tag}_cluster.log 2>&1
             summ ${tag}_cluster gather_cluster ;;
    bvh) timeout 600 $NCU -k regex:'prim_bounds|morton|leaf_records|karras|refit|collapse' -c 30 -o gpurun_out/${tag}_bvh python scripts/config_runs.py C1 > gpurun_out/${tag}_bvh.log 2>&1
         python scripts/ncu_summary.py gpurun_out/${tag}_bvh.ncu-rep > gpurun_out/${tag}_bvh_summary.txt 2>&1; rm -f gpurun_out/${tag}_bvh.ncu-rep ;;
    vsl) timeout 600 $NCU -k regex:gather_vsl -c 1 -o gpurun_out/${tag}_vsl python scripts/config_runs.py C3s > gpurun_out/${tag}_vsl.log 2>&1
         summ ${tag}_vsl gather_vsl; rm -f gpurun_out/${tag}_vsl.ncu-rep ;;
    lvc) timeout 600 $NCU -k regex:gather_lvc -c 1 -o gpurun_out/${tag}_lvc python scripts/config_runs.py LVC > gpurun_out/${tag}_lvc.log 2>&1
         summ ${tag}_lvc gather_lvc; rm -f gpurun_out/${tag}_lvc.ncu-rep ;;
    pt) timeout 600 $NCU -k regex:path_trace -c 1 -o gpurun_out/${tag}_pt python scripts/config_runs.py PT > gpurun_out/${tag}_pt.log 2>&1
        summ ${tag}_pt path_trace; rm -f gpurun_out/${tag}_pt.ncu-rep ;;
  esac
done
du -sh gpurun_out; tail -n 2 gpurun_out/${tag}_*.log | cut -c1-400
