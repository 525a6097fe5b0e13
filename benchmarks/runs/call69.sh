#!/bin/bash
O=gpurun_out/c69
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/smoke.log 2>&1
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:lowrank2_fused --csv --log-file $O/fused_launches.csv python bench.py --steps 1 --warmup 1 --no-msca --no-recon --no-lib --no-cpu > $O/ncu_fused.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/bench_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-lib > $O/ncu_list.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 python benchmarks/sweep_lowrank2.py --batches 1,16,256,1024 > $O/sweep.log 2>&1
echo done > $O/done
