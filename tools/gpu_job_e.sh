#!/bin/bash
# multi-GPU: the default bench line under torchrun (NCCL, graph-captured step with overlapped all-reduces)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_scale_$N.json 2> gpurun_out/r2_scale_$N.err; echo "torchrun N=$N rc=$?"
tail -c 1500 gpurun_out/r2_scale_$N.err
tail -c 300 gpurun_out/r2_scale_$N.json
