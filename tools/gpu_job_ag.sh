#!/bin/bash
# GPU job AG (2 GPUs): peer-memory all-reduce self-check against NCCL (correctness on all ranks + timing at three sizes)
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 tools/p2p_check.py > gpurun_out/r2_p2p_check_2gpu_final.txt 2>&1; echo "p2p_check rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2_p2p_check_2gpu_final.txt | tail -12
