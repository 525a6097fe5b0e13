#!/bin/bash
mkdir -p gpurun_out/c9
O=gpurun_out/c9
timeout 150 python -m pytest tests/test_lowrank2_fused_gpu.py tests/test_baseline_sizes_gpu.py -q --timeout 60 -x -k "not msca and not alexnet" > $O/pytest_fused.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $O/pytest_fused.log
if [ $rc -ne 0 ]; then echo "fused tests failed: skipping the rest" > $O/done; exit 0; fi
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for g in default 1; do
  if [ $g = default ]; then unset CNA_LR2_G1B; else export CNA_LR2_G1B=$g; fi
  timeout 300 python bench.py --per-layer --steps 10 --no-msca --no-recon --no-lib --no-cpu > $O/bench_g1b_$g.json 2> $O/bench_g1b_$g.err
done
unset CNA_LR2_G1B
timeout 300 python benchmarks/bench_dwconv.py --batch 256 > $O/dwconv_b256.log 2>&1
timeout 600 python benchmarks/low_rank_exp_spr.py > $O/spr_bf16.log 2>&1
timeout 600 python benchmarks/low_rank_exp_spr.py --decomp > $O/spr_bf16_decomp.log 2>&1
echo done > $O/done
