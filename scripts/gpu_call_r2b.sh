#!/bin/bash
# one GPU call: parity subset, headline bench, config-5 sweep, ncu of the cluster gather on a 1/16 tile share, region profile + histograms
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_baseline_configs.py tests/test_gpu_cluster.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 2 --no-cpu --no-single-frame > gpurun_out/b_r2b.json 2> gpurun_out/b_r2b.err
python scripts/photon_sweep.py 4194304,67108864 > gpurun_out/sweep_r2b_1gpu.jsonl 2> gpurun_out/sweep_r2b.err
scripts/ncu_r2.sh r2b cluster
EVPLP_LIB=$PWD/evplp_b200/lib/libevplp_b200_prof.so python bench.py --steps 1 --warmup 1 --no-cpu --no-single-frame --tile-share 4 > gpurun_out/prof_r2b.json 2> gpurun_out/prof_r2b.err
EVPLP_LIB=$PWD/evplp_b200/lib/libevplp_b200_hist.so python bench.py --steps 1 --warmup 1 --no-cpu --no-single-frame --tile-share 4 > gpurun_out/hist_r2b.json 2> gpurun_out/hist_r2b.err
tail -c 600 gpurun_out/b_r2b.err gpurun_out/sweep_r2b.err gpurun_out/prof_r2b.err; cat gpurun_out/sweep_r2b_1gpu.jsonl | cut -c1-500
