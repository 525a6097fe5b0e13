#!/bin/bash
O=gpurun_out/c44
mkdir -p $O
timeout 200 python bench.py --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_k64.json 2> $O/bench_k64.err
for k in 16 32; do
  timeout 200 python bench.py --rank $k --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_k${k}.json 2> $O/bench_k${k}.err
done
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python benchmarks/bench_msca.py --block > $O/msca_block.log 2>&1
timeout 300 python benchmarks/bench_lowrank1.py > $O/lowrank1.log 2>&1
echo done > $O/done
