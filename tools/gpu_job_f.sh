#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/p2p_check.py > gpurun_out/r2_p2p_check_$N.txt 2>&1; echo "p2p_check N=$N rc=$?"; grep -v OMP gpurun_out/r2_p2p_check_$N.txt | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_scale_p2p_$N.json 2> gpurun_out/r2_scale_p2p_$N.err; echo "torchrun N=$N rc=$?"
grep -v OMP gpurun_out/r2_scale_p2p_$N.err | tail -5
