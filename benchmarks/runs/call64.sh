#!/bin/bash
O=gpurun_out/c64
mkdir -p $O
for rep in 1 2; do
for v in main nobias; do
  if [ $v = nobias ]; then export CNA_B200_LIB=$PWD/benchmarks/alt_lib_nobias/libcna_b200.so; else unset CNA_B200_LIB; fi
  timeout 200 python bench.py --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_k64_${v}_$rep.json 2> $O/bench_k64_${v}_$rep.err
  timeout 200 python bench.py --rank 16 --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_k16_${v}_$rep.json 2> $O/bench_k16_${v}_$rep.err
done
done
echo done > $O/done
