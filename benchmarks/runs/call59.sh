#!/bin/bash
O=gpurun_out/c59
mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw_strip_tma --launch-skip 1 -c 1 -o $O/dw3x3_gelu -f python benchmarks/run_ffn_once.py --C 32 --exp 8 --H 128 > $O/ncu.log 2>&1
echo done > $O/done
