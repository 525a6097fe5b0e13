#!/bin/bash
O=gpurun_out/c62
mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu/recon_2rank.py > $O/recon_2rank.log 2>&1; echo "rc=$?" >> $O/recon_2rank.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "rc=$?" >> $O/bench_2gpu.err
echo done > $O/done
