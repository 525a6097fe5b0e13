#!/bin/bash
# scheme-1 host path (plan record, lean wrapper): every test that runs LowRankExpConvV1, smoke, the reference's num_base sweep
O=gpurun_out/c73
mkdir -p $O
S=$(date +%s)
timeout 40 python -m pytest tests/test_lowrank1_fused_gpu.py tests/test_modules_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q --timeout 60 -k "lowrank1 or lowrank_exp_v1 or empty or alexnet or fused_matches or basis or num_base or plan_record or full_size or random_shapes" > $O/pytest_v1.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - S ))" >> $O/pytest_v1.log
timeout 25 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/smoke.log 2>&1; echo "t=$(( $(date +%s) - S ))" >> $O/smoke.log
timeout 25 python benchmarks/low_rank_exp_spr.py --repeat 30 > $O/spr.jsonl 2> $O/spr.err; echo "rc=$? t=$(( $(date +%s) - S ))" >> $O/spr.err
if [ $(( $(date +%s) - S )) -lt 45 ]; then timeout 35 python -m pytest tests/test_runner_overlay_gpu.py -m gpu -q --timeout 60 -k "dodecomp or l2_reconstruct" > $O/pytest_runner.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - S ))" >> $O/pytest_runner.log; fi
echo done > $O/done
