#!/bin/bash
mkdir -p gpurun_out/c12
O=gpurun_out/c12
# reconstruction step at the per-GPU batch of the 8-GPU run (32 images) on one GPU: which kernels do not shrink with the batch?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lowrank2|wgrad|l2_|recon_sync|tapconv" -c 420 --csv --log-file $O/recon_b32_launches.csv python bench.py --steps 3 --no-msca --no-lib --no-cpu --recon-batch 32 > $O/recon_b32_ncu.log 2>&1
timeout 300 python bench.py --steps 10 --no-msca --no-lib --no-cpu --recon-batch 32 > $O/bench_b32.json 2> $O/bench_b32.err
echo done > $O/done
