#!/bin/bash
mkdir -p gpurun_out/c8
O=gpurun_out/c8
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
echo done > $O/done
