#!/bin/bash
O=gpurun_out/c58
mkdir -p $O
timeout 600 python -m pytest tests/test_mscan_block_gpu.py tests/test_mscan_attention_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 200 > $O/pytest_a.log 2>&1; echo "rc=$?" >> $O/pytest_a.log
timeout 600 python benchmarks/bench_msca.py --block > $O/msca_block.log 2>&1
timeout 600 python bench.py --steps 5 --no-recon --no-lib --no-cpu > $O/bench.json 2> $O/bench.err
echo done > $O/done
