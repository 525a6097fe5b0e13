#!/bin/bash
O=gpurun_out/c68
mkdir -p $O
timeout 900 python -m pytest tests/test_lowrank2_fused_gpu.py tests/test_baseline_sizes_gpu.py tests/test_recon_gpu.py -m gpu -q -x --timeout 250 > $O/pytest_a.log 2>&1; echo "rc=$?" >> $O/pytest_a.log
for rep in 1 2; do
for v in new prev; do
  unset CNA_B200_LIB
  if [ $v = prev ]; then export CNA_B200_LIB=$PWD/benchmarks/alt_lib_prev/libcna_b200.so; fi
  timeout 200 python bench.py --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_k64_${v}_$rep.json 2> $O/bench_k64_${v}_$rep.err
  timeout 200 python bench.py --rank 16 --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_k16_${v}_$rep.json 2> $O/bench_k16_${v}_$rep.err
  timeout 200 python bench.py --rank 32 --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_k32_${v}_$rep.json 2> $O/bench_k32_${v}_$rep.err
done
done
echo done > $O/done
