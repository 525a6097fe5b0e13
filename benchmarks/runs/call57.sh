#!/bin/bash
O=gpurun_out/c57
mkdir -p $O
for cfg in "32 8 128" "64 8 64" "160 4 32" "256 4 16"; do
  set -- $cfg
  timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv1x1_tc_kernel|dw_strip_tma" --launch-skip 0 --csv --log-file $O/ffn_$1.csv python benchmarks/run_ffn_once.py --C $1 --exp $2 --H $3 > $O/ffn_$1.log 2>&1
done
echo done > $O/done
