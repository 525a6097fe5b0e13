#!/bin/bash
mkdir -p gpurun_out/c27
O=gpurun_out/c27
timeout 300 python -m pytest tests/test_modules_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q --timeout 100 -x > $O/pytest_modules.log 2>&1; echo "pytest rc=$?" >> $O/pytest_modules.log
timeout 200 python benchmarks/bench_msca.py > $O/msca.log 2>&1; echo "rc=$?" >> $O/msca.log
timeout 120 python benchmarks/bench_mixgate.py > $O/mixgate.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"msca_sd|conv1x1_tc|dw_tile" --csv --log-file $O/msca_launches.csv python benchmarks/bench_msca.py --iters 1 > $O/ncu_list.log 2>&1
echo done > $O/done
