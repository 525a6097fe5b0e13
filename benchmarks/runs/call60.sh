#!/bin/bash
O=gpurun_out/c60
mkdir -p $O
timeout 600 python -m pytest tests/test_mscan_block_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 200 > $O/pytest_a.log 2>&1; echo "rc=$?" >> $O/pytest_a.log
timeout 300 python benchmarks/bench_dwconv.py --batch 256 --iters 10 > $O/dwconv_b256.log 2>&1
timeout 600 python benchmarks/bench_msca.py --block > $O/msca_block.log 2>&1
echo done > $O/done
