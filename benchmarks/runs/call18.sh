#!/bin/bash
mkdir -p gpurun_out/c18
O=gpurun_out/c18
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q --timeout 200 -k "msca or mscan" > $O/pytest_msca.log 2>&1; echo "pytest rc=$?" >> $O/pytest_msca.log
timeout 300 python benchmarks/bench_msca.py > $O/msca_mma.log 2>&1
CNA_MSCA_SD=f timeout 300 python benchmarks/bench_msca.py > $O/msca_ffma.log 2>&1
echo done > $O/done
