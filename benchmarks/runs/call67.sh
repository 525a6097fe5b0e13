#!/bin/bash
O=gpurun_out/c67
mkdir -p $O
timeout 900 python -m pytest tests/test_lowrank2_fused_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q -x --timeout 250 > $O/pytest_a.log 2>&1; echo "rc=$?" >> $O/pytest_a.log
CNA_LR2_VERBOSE=1 timeout 200 python bench.py --per-layer --steps 2 --no-msca --no-recon --no-lib --no-cpu 2>&1 | grep "plan:" | sort | uniq > $O/plans.log
for rep in 1 2; do
for v in new off prev; do
  unset CNA_B200_LIB; unset CNA_LR2_BIASMMA
  if [ $v = prev ]; then export CNA_B200_LIB=$PWD/benchmarks/alt_lib_prev/libcna_b200.so; fi
  if [ $v = off ]; then export CNA_LR2_BIASMMA=0; fi
  timeout 200 python bench.py --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_k64_${v}_$rep.json 2> $O/bench_k64_${v}_$rep.err
  timeout 200 python bench.py --rank 16 --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_k16_${v}_$rep.json 2> $O/bench_k16_${v}_$rep.err
done
done
echo done > $O/done
