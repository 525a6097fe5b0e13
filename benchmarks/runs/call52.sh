#!/bin/bash
O=gpurun_out/c52
mkdir -p $O
timeout 600 python -m pytest tests/test_lowrank2_fused_gpu.py tests/test_baseline_sizes_gpu.py tests/test_recon_gpu.py -m gpu -q -x --timeout 250 > $O/pytest_a.log 2>&1; echo "rc=$?" >> $O/pytest_a.log
CNA_B200_LIB=$PWD/benchmarks/alt_lib_prev/libcna_b200.so timeout 200 python bench.py --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_prev.json 2> $O/bench_prev.err
CNA_LR2_VERBOSE=1 timeout 200 python bench.py --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_auto.json 2> $O/bench_auto.err
for m in 01 10 00 11; do
  CNA_LR2_VERBOSE=1 CNA_LR2_MODE=$m timeout 200 python bench.py --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_m$m.json 2> $O/bench_m$m.err
done
CNA_B200_LIB=$PWD/benchmarks/alt_lib_prev/libcna_b200.so timeout 200 python bench.py --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_prev2.json 2> $O/bench_prev2.err
echo done > $O/done
