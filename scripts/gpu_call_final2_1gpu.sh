#!/bin/bash
# round-2 closing single-GPU evidence (after the adaptive cluster layout): full -m gpu suite, default bench line, DRAM traffic of the
# headline gather launch and of the splat kernels, --set full of the headline kernel on a 1/16 tile share
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r2_final_1gpu.json 2> gpurun_out/bench_r2_final_1gpu.err
cut -c1-160 gpurun_out/bench_r2_final_1gpu.json
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'gather_cluster|splat_prepare|splat_fill|splat_tile' -c 4 --csv --log-file gpurun_out/r2_traffic_headline_final.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-single-frame > /dev/null 2>&1
grep -c . gpurun_out/r2_traffic_headline_final.csv
scripts/ncu_r2.sh r2final cluster > /dev/null 2>&1
python scripts/ncu_lines.py gpurun_out/r2final_cluster.ncu-rep 0 "tri_test=device_scene.h:96-117,gather_fast.cu:326-345;ray slab of a candidate=device_scene.h:160-195,gather_fast.cu:313-325;dshaft_overlap=gather_fast.cu:210-221;make_dshaft=gather_fast.cu:185-208;store/load shaft + tile box + texel=gather_fast.cu:223-223,gather_fast.cu:233-244,gather_fast.cu:282-291,gather_fast.cu:351-357;descent loop=gather_fast.cu:245-276;filter loop=gather_fast.cu:292-312,gather_fast.cu:346-349;stage+live=gather_fast.cu:464-497;depth groups+tile setup=gather_fast.cu:402-463;shared driver=gather_fast.cu:498-538;perVPL driver=gather_fast.cu:539-579;shading=gather_fast.cu:580-652,gather_fast.cu:47-73" >> gpurun_out/r2final_cluster_summary.txt 2>&1
rm -f gpurun_out/r2final_cluster.ncu-rep
