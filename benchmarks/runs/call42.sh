#!/bin/bash
O=gpurun_out/c42
mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lowrank2_fused --launch-skip 2 -c 1 -o $O/l1_full -f python benchmarks/run_one_layer.py --C 3 --N 64 --K 64 --H 224 --batch 256 --iters 3 > $O/ncu_l1.log 2>&1
echo done > $O/done
