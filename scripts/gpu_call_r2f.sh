#!/bin/bash
NCU="ncu --set full --clock-control none --import-source on -f"
summ() { python scripts/ncu_summary.py gpurun_out/$1.ncu-rep "$2" > gpurun_out/$1_summary.txt 2>&1; python scripts/ncu_hot.py gpurun_out/$1.ncu-rep >> gpurun_out/$1_summary.txt 2>&1; python scripts/ncu_lines.py gpurun_out/$1.ncu-rep 30 >> gpurun_out/$1_summary.txt 2>&1; }
timeout 300 $NCU -k regex:light_trace -c 1 -o gpurun_out/r2f_lt python scripts/photon_sweep.py 4194304 > gpurun_out/r2f_lt.log 2>&1; summ r2f_lt light_trace; rm -f gpurun_out/r2f_lt.ncu-rep
timeout 300 $NCU -k regex:splat_tile -c 1 -o gpurun_out/r2f_st python scripts/photon_sweep.py 4194304 > gpurun_out/r2f_st.log 2>&1; summ r2f_st splat_tile; rm -f gpurun_out/r2f_st.ncu-rep
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r2_v5_1gpu.json 2> gpurun_out/bench_r2_v5_1gpu.err
cut -c1-200 gpurun_out/bench_r2_v5_1gpu.json
