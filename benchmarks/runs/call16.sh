#!/bin/bash
mkdir -p gpurun_out/c16
O=gpurun_out/c16
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python benchmarks/bench_dwconv.py --batch 256 > $O/dwconv_b256.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dw_ --csv --log-file $O/dwconv_b256_ncu.csv python benchmarks/bench_dwconv.py --batch 256 --iters 1 > $O/dwconv_ncu.log 2>&1
timeout 600 python benchmarks/low_rank_exp_spr.py > $O/spr_bf16.log 2>&1
echo done > $O/done
