#!/bin/bash
# scheme-1 16-basis instance: its tests first, the reference's num_base sweep, then the whole GPU suite, smoke, and the bench line
O=gpurun_out/c71
mkdir -p $O
S=$(date +%s)
timeout 90 python -m pytest tests/test_lowrank1_fused_gpu.py -m gpu -q --timeout 60 > $O/pytest_l1.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - S ))" >> $O/pytest_l1.log
timeout 60 python benchmarks/low_rank_exp_spr.py --repeat 30 > $O/spr.jsonl 2> $O/spr.err; echo "rc=$? t=$(( $(date +%s) - S ))" >> $O/spr.err
timeout 170 python -m pytest tests -m gpu -q --timeout 120 --deselect tests/test_lowrank1_fused_gpu.py > $O/pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - S ))" >> $O/pytest.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/smoke.log 2>&1
if [ $(( $(date +%s) - S )) -lt 200 ]; then timeout 60 python bench.py --no-cpu > $O/bench.json 2> $O/bench.err; echo "rc=$? t=$(( $(date +%s) - S ))" >> $O/bench.err; fi
echo done > $O/done
