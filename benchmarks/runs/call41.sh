#!/bin/bash
O=gpurun_out/c41
mkdir -p $O
timeout 100 python benchmarks/trace_lowrank2.py --C 3 --N 64 --K 64 --H 224 --batch 256 --tiles 48 --quiet 1 > $O/trace_l1_y3.log 2>&1
CNA_LR2_YSTG=1 timeout 100 python benchmarks/trace_lowrank2.py --C 3 --N 64 --K 64 --H 224 --batch 256 --tiles 48 --quiet 1 > $O/trace_l1_y1.log 2>&1
for d in 0 1 2 3 64 128; do
  echo "debug=$d" >> $O/l1_debug.log
  CNA_LR2_DEBUG=$d timeout 100 python benchmarks/run_one_layer.py --C 3 --N 64 --K 64 --H 224 --batch 256 --iters 5 >> $O/l1_debug.log 2>&1
done
for d in 0 1 3; do
  echo "L3 debug=$d" >> $O/l1_debug.log
  CNA_LR2_DEBUG=$d timeout 100 python benchmarks/run_one_layer.py --C 64 --N 128 --K 64 --H 112 --batch 256 --iters 5 >> $O/l1_debug.log 2>&1
done
echo done > $O/done
