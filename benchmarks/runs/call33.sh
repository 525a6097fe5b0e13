#!/bin/bash
mkdir -p gpurun_out/c33
O=gpurun_out/c33
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "rc=$?" >> $O/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_2gpu.json 2> $O/bench_ref_2gpu.err; echo "rc=$?" >> $O/bench_ref_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tests/multi_gpu/recon_2rank.py > $O/recon_2rank.log 2>&1; echo "rc=$?" >> $O/recon_2rank.log
echo done > $O/done
