#!/bin/bash
mkdir -p gpurun_out/c32
O=gpurun_out/c32
timeout 300 python -m pytest tests/test_mscan_attention_gpu.py tests/test_modules_gpu.py -m gpu -q --timeout 100 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python benchmarks/bench_msca.py --block > $O/msca_block.log 2>&1; echo "rc=$?" >> $O/msca_block.log
echo done > $O/done
