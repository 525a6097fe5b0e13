#!/bin/bash
# 8 GPUs: reconstruction strong scaling (fused peer-memory exchange kernel), inference / e2e scaling with uint8 input
mkdir -p gpurun_out/c11
O=gpurun_out/c11
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu/recon_2rank.py > $O/recon_8rank.log 2>&1; echo "rc=$?" >> $O/recon_8rank.log
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 10 --warmup 3 > $O/bench$n.json 2> $O/bench$n.err; echo "rc=$?" >> $O/bench$n.err
done
timeout 300 python bench.py --steps 10 --no-lib --no-cpu > $O/bench1.json 2> $O/bench1.err
echo done > $O/done
