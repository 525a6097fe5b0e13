#!/bin/bash
mkdir -p gpurun_out/c2
O=gpurun_out/c2
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 900 python bench.py --steps 10 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
echo done > $O/done
