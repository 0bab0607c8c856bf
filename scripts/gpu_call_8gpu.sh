#!/bin/bash
# 8-GPU measurements: headline bench (weak scaling, iterations round-robin; same-job checksums; single-frame strong scaling) and the
# config-5 photon sweep partitioned by light paths
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench_r2_final_8gpu.json 2> gpurun_out/bench_r2_final_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/photon_sweep.py 67108864,268435456 > gpurun_out/sweep_r2_final_8gpu.jsonl 2> gpurun_out/sweep_r2_final_8gpu.err
tail -c 300 gpurun_out/bench_r2_final_8gpu.err; cut -c1-300 gpurun_out/bench_r2_final_8gpu.json; cut -c1-420 gpurun_out/sweep_r2_final_8gpu.jsonl; tail -c 300 gpurun_out/sweep_r2_final_8gpu.err
