#!/bin/bash
mkdir -p gpurun_out/c36
O=gpurun_out/c36
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "rc=$?" >> $O/bench_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tests/multi_gpu/recon_2rank.py > $O/recon_8rank.log 2>&1; echo "rc=$?" >> $O/recon_8rank.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 20 --warmup 5 --no-msca > $O/bench_4gpu.json 2> $O/bench_4gpu.err; echo "rc=$?" >> $O/bench_4gpu.err
echo done > $O/done
