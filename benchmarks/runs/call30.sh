#!/bin/bash
mkdir -p gpurun_out/c30
O=gpurun_out/c30
CNA_MSCA_WS=1 CNA_MSCA_NIO=8 timeout 60 python -m pytest tests/test_modules_gpu.py -m gpu -q --timeout 50 -x -k "test_msca_fused_sd_kernel_vs_oracle and dt1-32-40-50" > $O/pytest_one.log 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/pytest_one.log
if [ $rc -ne 0 ]; then echo done > $O/done; exit 0; fi
CNA_MSCA_WS=1 CNA_MSCA_NIO=8 timeout 300 python -m pytest tests/test_modules_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q --timeout 100 -x -k "msca or mscan" > $O/pytest_msca.log 2>&1; echo "pytest rc=$?" >> $O/pytest_msca.log
CNA_MSCA_WS=1 CNA_MSCA_NIO=8 timeout 200 python benchmarks/bench_msca.py > $O/msca_ws8.log 2>&1; echo "rc=$?" >> $O/msca_ws8.log
CNA_MSCA_WS=1 CNA_MSCA_NIO=4 timeout 200 python benchmarks/bench_msca.py > $O/msca_ws4.log 2>&1; echo "rc=$?" >> $O/msca_ws4.log
echo done > $O/done
