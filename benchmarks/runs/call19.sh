#!/bin/bash
mkdir -p gpurun_out/c19
O=gpurun_out/c19
timeout 300 python benchmarks/bench_msca.py > $O/msca_mma.log 2>&1
CNA_MSCA_SD=f timeout 300 python benchmarks/bench_msca.py > $O/msca_ffma.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"msca_sdm|conv1x1|dw_tile" --csv --log-file $O/msca_launches.csv python benchmarks/bench_msca.py --iters 1 > $O/ncu_list.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:msca_sdm -s 3 -c 1 -o $O/msca_sdm python benchmarks/bench_msca.py --iters 1 > $O/ncu_full.log 2>&1
echo done > $O/done
