#!/bin/bash
mkdir -p gpurun_out/c26
O=gpurun_out/c26
timeout 60 python -m pytest tests/test_modules_gpu.py -m gpu -q --timeout 50 -x -k "test_msca_fused_sd_kernel_vs_oracle and dt1-32-40-50" > $O/pytest_one.log 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/pytest_one.log
if [ $rc -ne 0 ]; then echo done > $O/done; exit 0; fi
timeout 300 python -m pytest tests/test_modules_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q --timeout 100 -x -k "msca or mscan" > $O/pytest_msca.log 2>&1; echo "pytest rc=$?" >> $O/pytest_msca.log
timeout 200 python benchmarks/bench_msca.py > $O/msca_ws.log 2>&1; echo "rc=$?" >> $O/msca_ws.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:msca_sdw -s 3 -c 1 -o $O/msca_sdw python benchmarks/bench_msca.py --iters 1 > $O/ncu_full.log 2>&1
echo done > $O/done
CNA_MSCA_CPW=2 timeout 200 python benchmarks/bench_msca.py > gpurun_out/c26/msca_cpw2.log 2>&1
