#!/bin/bash
O=gpurun_out/c55
mkdir -p $O
timeout 600 python -m pytest tests/test_mscan_block_gpu.py tests/test_mscan_attention_gpu.py tests/test_kernels_gpu.py -m gpu -q -x --timeout 200 > $O/pytest_a.log 2>&1; echo "rc=$?" >> $O/pytest_a.log
echo done > $O/done
