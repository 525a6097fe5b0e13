#!/bin/bash
O=gpurun_out/c70
mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "rc=$?" >> $O/bench_8gpu.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu/recon_2rank.py > $O/recon_8rank.log 2>&1; echo "rc=$?" >> $O/recon_8rank.log
echo done > $O/done
