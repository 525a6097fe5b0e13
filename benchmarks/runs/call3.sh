#!/bin/bash
mkdir -p gpurun_out/c3
O=gpurun_out/c3
timeout 600 python -m pytest tests -m gpu -q -k "msca or smoke or MSCA or mscan" > $O/pytest_msca.log 2>&1; echo "pytest rc=$?" >> $O/pytest_msca.log
timeout 300 python benchmarks/bench_msca.py > $O/msca_new.log 2>&1
CNA_MSCA_SD=old timeout 300 python benchmarks/bench_msca.py > $O/msca_old.log 2>&1
timeout 300 python benchmarks/bench_msca.py --batch 256 > $O/msca_new_b256.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lowrank2|wgrad|l2_|recon_sync|tapconv" -c 400 --csv --log-file $O/recon_launches.csv python bench.py --steps 3 --no-msca --no-lib --no-cpu > $O/recon_ncu.log 2>&1
echo done > $O/done
