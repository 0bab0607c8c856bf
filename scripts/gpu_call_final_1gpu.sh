#!/bin/bash
# round-2 final single-GPU evidence: full -m gpu suite, the default bench line, launch list, --set full of the headline kernel on a
# 1/16 tile share, DRAM traffic of the headline launch, config-5 sweep
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r2_final_1gpu.json 2> gpurun_out/bench_r2_final_1gpu.err
cut -c1-160 gpurun_out/bench_r2_final_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_default_final.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-single-frame > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_bench_default_final.csv > gpurun_out/r2_launches_bench_default_final_summary.txt 2>&1; head -8 gpurun_out/r2_launches_bench_default_final_summary.txt
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gather_cluster -c 1 --csv --log-file gpurun_out/r2_gather_traffic_headline_final.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-single-frame > /dev/null 2>&1
tail -4 gpurun_out/r2_gather_traffic_headline_final.csv
scripts/ncu_r2.sh r2final cluster > /dev/null 2>&1
python scripts/ncu_lines.py gpurun_out/r2final_cluster.ncu-rep 0 "tri_test=device_scene.h:96-117,gather_fast.cu:272-293;ray slab of a candidate=device_scene.h:160-195,gather_fast.cu:258-271;dshaft_overlap=gather_fast.cu:166-177;make_dshaft=gather_fast.cu:141-164;store/load shaft=gather_fast.cu:179-179,gather_fast.cu:189-200,gather_fast.cu:296-302;descent loop=gather_fast.cu:201-232;filter loop=gather_fast.cu:237-257,gather_fast.cu:294-295;stage+live=gather_fast.cu:402-420;depth groups+tile setup=gather_fast.cu:347-401;shared driver=gather_fast.cu:421-470;perVPL driver=gather_fast.cu:471-513;shading=gather_fast.cu:514-585,gather_fast.cu:46-70" >> gpurun_out/r2final_cluster_summary.txt 2>&1
rm -f gpurun_out/r2final_cluster.ncu-rep
python scripts/photon_sweep.py 1048576,4194304,67108864 2>/dev/null > gpurun_out/sweep_r2_final_1gpu.jsonl; cut -c1-300 gpurun_out/sweep_r2_final_1gpu.jsonl
