#!/bin/bash
mkdir -p gpurun_out/c34
O=gpurun_out/c34
timeout 600 python -m pytest tests/test_modules_gpu.py -m gpu -q --timeout 250 -k "generations" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
echo done > $O/done
