#!/bin/bash
mkdir -p gpurun_out/c29
O=gpurun_out/c29
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
echo done > $O/done
