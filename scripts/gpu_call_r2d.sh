#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_cluster.py tests/test_gpu_fullsize.py tests/test_gpu_baseline_configs.py -m gpu -x -q 2>&1 | tail -3
scripts/quick_variants.sh "--no-single-frame" ""
cp gpurun_out/q.json gpurun_out/b_r2d.json
python scripts/photon_sweep.py 4194304,67108864 > gpurun_out/sweep_r2d_1gpu.jsonl 2> gpurun_out/sweep_r2d.err; cut -c1-420 gpurun_out/sweep_r2d_1gpu.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/sweep_launches_r2d.csv python scripts/photon_sweep.py 4194304 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/sweep_launches_r2d.csv 2>&1 | tail -15
EVPLP_LIB=$PWD/evplp_b200/lib/libevplp_b200_hist.so python bench.py --steps 1 --warmup 1 --no-cpu --no-single-frame --tile-share 4 > gpurun_out/hist_r2d.json 2> gpurun_out/hist_r2d.err
