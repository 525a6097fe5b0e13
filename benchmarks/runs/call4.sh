#!/bin/bash
mkdir -p gpurun_out/c4
O=gpurun_out/c4
timeout 600 python -m pytest tests -m gpu -q -k "msca or smoke or MSCA or mscan" > $O/pytest_msca.log 2>&1; echo "pytest rc=$?" >> $O/pytest_msca.log
timeout 300 python benchmarks/bench_msca.py > $O/msca_new.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:msca_sd16 -s 3 -c 1 -o $O/msca_sd16 python benchmarks/bench_msca.py --iters 1 > $O/ncu.log 2>&1
echo done > $O/done
