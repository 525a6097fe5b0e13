#!/bin/bash
mkdir -p gpurun_out/c10
O=gpurun_out/c10
# A/B of the fused conv kernel revisions on ONE box: r2a = the kernel before the second GEMM1 issuer / merged producer
for rep in 1 2; do
  CNA_B200_LIB=$PWD/convnet-approximater_b200/lib/libcna_b200_r2a.so timeout 300 python bench.py --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_old_$rep.json 2> $O/bench_old_$rep.err
  timeout 300 python bench.py --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_new_$rep.json 2> $O/bench_new_$rep.err
  CNA_LR2_G1B=0 timeout 300 python bench.py --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_new_g0_$rep.json 2> $O/bench_new_g0_$rep.err
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dw_strip -s 2 -c 1 -o $O/strip_1x7 python benchmarks/run_strip_once.py > $O/ncu_strip.log 2>&1
timeout 600 python -m pytest tests/test_runner_overlay_gpu.py -q --timeout 300 > $O/pytest_overlay.log 2>&1; echo "rc=$?" >> $O/pytest_overlay.log
echo done > $O/done
