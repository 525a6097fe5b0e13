#!/bin/bash
O=gpurun_out/c43
mkdir -p $O
timeout 600 python -m pytest tests/test_lowrank2_fused_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q -x --timeout 250 > $O/pytest_a.log 2>&1; echo "rc=$?" >> $O/pytest_a.log
timeout 200 python bench.py --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_k64.json 2> $O/bench_k64.err
for k in 16 32; do
  timeout 200 python bench.py --rank $k --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_k${k}.json 2> $O/bench_k${k}.err
done
timeout 100 python benchmarks/trace_lowrank2.py --C 3 --N 64 --K 64 --H 224 --batch 256 --tiles 48 --quiet 1 > $O/trace_l1.log 2>&1
echo done > $O/done
