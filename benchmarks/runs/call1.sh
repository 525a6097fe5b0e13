#!/bin/bash
# GPU call 1 of round 2: parity suite, bench at HEAD, per-tile traces of the fused conv, DRAM counters of the depthwise kernels
mkdir -p gpurun_out/c1
O=gpurun_out/c1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python bench.py --per-layer --steps 10 > $O/bench.json 2> $O/bench.err
for spec in "3 64 64 224 64" "64 64 64 224 64" "64 64 16 224 64" "64 128 64 112 128" "256 256 64 56 256" "512 512 64 28 256" "512 512 64 14 256"; do
  set -- $spec
  timeout 300 python benchmarks/trace_lowrank2.py --C $1 --N $2 --K $3 --H $4 --batch $5 --tiles 48 --quiet 1 > $O/trace_C$1_N$2_K$3_H$4.log 2>&1
done
timeout 600 python benchmarks/bench_msca.py --torch > $O/msca.log 2>&1
timeout 600 python benchmarks/bench_dwconv.py --batch 256 > $O/dwconv_b256.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dw_ --csv --log-file $O/dwconv_b256_ncu.csv python benchmarks/bench_dwconv.py --batch 256 --iters 1 > $O/dwconv_ncu.log 2>&1
echo done > $O/done
