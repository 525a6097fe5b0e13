#!/bin/bash
mkdir -p gpurun_out/c15
O=gpurun_out/c15
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --steps 10 --warmup 3 --no-msca > $O/bench$n.json 2> $O/bench$n.err; echo "rc=$?" >> $O/bench$n.err
done
timeout 300 python bench.py --steps 10 --no-lib --no-cpu --no-msca > $O/bench1.json 2> $O/bench1.err
echo done > $O/done
