#!/bin/bash
mkdir -p gpurun_out/c20
O=gpurun_out/c20
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q --timeout 200 -k "msca or mscan" > $O/pytest_msca.log 2>&1; echo "pytest rc=$?" >> $O/pytest_msca.log
timeout 300 python benchmarks/bench_msca.py > $O/msca_cpw2.log 2>&1
CNA_MSCA_CPW=1 timeout 300 python benchmarks/bench_msca.py > $O/msca_cpw1.log 2>&1
CNA_MSCA_CPW=1 timeout 600 python -m pytest tests/test_modules_gpu.py -m gpu -q --timeout 200 -k "msca" > $O/pytest_cpw1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:msca_sdm -s 3 -c 1 -o $O/msca_sdm python benchmarks/bench_msca.py --iters 1 > $O/ncu_full.log 2>&1
echo done > $O/done
