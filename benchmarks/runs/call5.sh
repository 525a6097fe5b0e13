#!/bin/bash
# 2 GPUs: fused reduce-scatter + AdamW + all-gather kernel over peer memory, then the bench with its recon sub-record
mkdir -p gpurun_out/c5
O=gpurun_out/c5
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu/recon_2rank.py > $O/recon_2rank.log 2>&1; echo "rc=$?" >> $O/recon_2rank.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench2.json 2> $O/bench2.err; echo "rc=$?" >> $O/bench2.err
CNA_MSCA_FUSED=1 timeout 300 python benchmarks/bench_msca.py > $O/msca_fused1.log 2>&1
echo done > $O/done
