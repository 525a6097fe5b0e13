#!/bin/bash
O=gpurun_out/c49
mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw_strip_tma --launch-skip 1 -c 1 -o $O/strip_1x21 -f python benchmarks/run_strip_once.py --kh 1 --kw 21 > $O/ncu_1x21.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw_strip_tma --launch-skip 1 -c 1 -o $O/strip_1x7 -f python benchmarks/run_strip_once.py --kh 1 --kw 7 > $O/ncu_1x7.log 2>&1
echo done > $O/done
