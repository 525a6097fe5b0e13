#!/bin/bash
O=gpurun_out/c54
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -m gpu -q -x --timeout 200 > $O/pytest_a.log 2>&1; echo "rc=$?" >> $O/pytest_a.log
timeout 300 python benchmarks/bench_dwconv.py --batch 256 --iters 10 > $O/dwconv_b256.log 2>&1
CNA_DW_NO_TMA_SQUARE=1 timeout 300 python benchmarks/bench_dwconv.py --batch 256 --iters 10 > $O/dwconv_b256_nosq.log 2>&1
echo done > $O/done
