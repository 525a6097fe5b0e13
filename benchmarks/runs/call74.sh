#!/bin/bash
# the layers whose forward() gained the empty-batch wrapper and were not in call 72 / 73: MSCAN attention / block, VGG trunk module path
O=gpurun_out/c74
mkdir -p $O
timeout 17 python -m pytest tests/test_mscan_attention_gpu.py tests/test_mscan_block_gpu.py tests/test_lowrank2_fused_gpu.py -m gpu -q -x --timeout 30 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
echo done > $O/done
