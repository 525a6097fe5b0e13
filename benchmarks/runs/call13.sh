#!/bin/bash
mkdir -p gpurun_out/c13
O=gpurun_out/c13
timeout 600 python -m pytest tests/test_recon_gpu.py -q --timeout 120 > $O/pytest_recon.log 2>&1; echo "rc=$?" >> $O/pytest_recon.log
timeout 300 python bench.py --steps 10 --no-msca --no-lib --no-cpu --recon-batch 32 > $O/bench_b32.json 2> $O/bench_b32.err
timeout 300 python bench.py --steps 10 --no-msca --no-lib --no-cpu > $O/bench_b256.json 2> $O/bench_b256.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lowrank2|wgrad|l2_|recon_sync|tapconv" -c 420 --csv --log-file $O/recon_b32_launches.csv python bench.py --steps 3 --no-msca --no-lib --no-cpu --recon-batch 32 > $O/recon_b32_ncu.log 2>&1
echo done > $O/done
