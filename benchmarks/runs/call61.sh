#!/bin/bash
O=gpurun_out/c61
mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 200 -k "dwconv or tma_strip" > $O/pytest_a.log 2>&1; echo "rc=$?" >> $O/pytest_a.log
for rep in 1 2; do
timeout 300 python benchmarks/bench_dwconv.py --batch 256 --iters 10 > $O/dw_main_$rep.log 2>&1
CNA_B200_LIB=$PWD/benchmarks/alt_lib_ahead/libcna_b200.so timeout 300 python benchmarks/bench_dwconv.py --batch 256 --iters 10 > $O/dw_ahead_$rep.log 2>&1
done
echo done > $O/done
