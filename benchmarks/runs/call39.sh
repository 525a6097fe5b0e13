#!/bin/bash
O=gpurun_out/c39
mkdir -p $O
timeout 600 python -m pytest tests/test_lowrank2_fused_gpu.py tests/test_mscan_attention_gpu.py -m gpu -q -x --timeout 250 > $O/pytest_a.log 2>&1; echo "rc=$?" >> $O/pytest_a.log
timeout 600 python -m pytest tests/test_baseline_sizes_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 250 > $O/pytest_b.log 2>&1; echo "rc=$?" >> $O/pytest_b.log
for y in 1 2 3 4 auto; do
  if [ $y = auto ]; then unset CNA_LR2_YSTG; else export CNA_LR2_YSTG=$y; fi
  timeout 200 python bench.py --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_y$y.json 2> $O/bench_y$y.err
done
unset CNA_LR2_YSTG
for k in 16 32; do
  timeout 200 python bench.py --rank $k --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_k${k}.json 2> $O/bench_k${k}.err
  CNA_LR2_YSTG=1 timeout 200 python bench.py --rank $k --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_k${k}_y1.json 2> $O/bench_k${k}_y1.err
done
timeout 200 python benchmarks/bench_mixgate.py > $O/mixgate.log 2>&1
timeout 300 python benchmarks/bench_msca.py --block > $O/msca_block.log 2>&1
echo done > $O/done
