#!/bin/bash
O=gpurun_out/c63
mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 600 -k "tma_strip or dwconv3x3 or (dwconv_vs_torch and bfloat16)" > $O/memcheck_dw.log 2>&1; echo "rc=$?" >> $O/memcheck_dw.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_mscan_block_gpu.py -m gpu -q -x --timeout 600 > $O/memcheck_block.log 2>&1; echo "rc=$?" >> $O/memcheck_block.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_lowrank2_fused_gpu.py -m gpu -q -x --timeout 600 -k "not full_size" > $O/memcheck_lr2.log 2>&1; echo "rc=$?" >> $O/memcheck_lr2.log
echo done > $O/done
