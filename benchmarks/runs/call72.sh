#!/bin/bash
# scheme-1 2-basis instance + empty-batch handling: the scheme-1 tests, the module tests, the reference's num_base sweep
O=gpurun_out/c72
mkdir -p $O
S=$(date +%s)
timeout 60 python -m pytest tests/test_lowrank1_fused_gpu.py -m gpu -q --timeout 60 > $O/pytest_l1.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - S ))" >> $O/pytest_l1.log
timeout 40 python benchmarks/low_rank_exp_spr.py --repeat 30 > $O/spr.jsonl 2> $O/spr.err; echo "rc=$? t=$(( $(date +%s) - S ))" >> $O/spr.err
timeout 70 python -m pytest tests/test_modules_gpu.py -m gpu -v --timeout 60 > $O/pytest_mod.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - S ))" >> $O/pytest_mod.log
echo done > $O/done
