#!/bin/bash
mkdir -p gpurun_out/c31
O=gpurun_out/c31
timeout 300 python -m pytest tests/test_modules_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q --timeout 100 -x -k "msca or mscan" > $O/pytest_msca.log 2>&1; echo "pytest rc=$?" >> $O/pytest_msca.log
timeout 200 python benchmarks/bench_msca.py > $O/msca.log 2>&1; echo "rc=$?" >> $O/msca.log
echo done > $O/done
