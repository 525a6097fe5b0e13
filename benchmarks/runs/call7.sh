#!/bin/bash
mkdir -p gpurun_out/c7
O=gpurun_out/c7
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python bench.py --steps 10 --no-lib --no-cpu > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
echo done > $O/done
