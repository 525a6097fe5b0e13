#!/bin/bash
O=gpurun_out/c56
mkdir -p $O
timeout 600 python benchmarks/bench_msca.py --block > $O/msca_block.log 2>&1
echo done > $O/done
