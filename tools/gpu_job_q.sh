#!/bin/bash
# GPU job Q (2 GPUs): one-launch peer-memory all-reduce: self-check against NCCL, then config 3
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/p2p_check.py > gpurun_out/r2_p2p_check_2gpu.txt 2>&1; echo "p2p_check rc=$?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2_p2p_check_2gpu.txt | tail -18
grep -q "P2P_CHECK OK" gpurun_out/r2_p2p_check_2gpu.txt || exit 1
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --config c3 --steps 100 --warmup 5 --no-cpu-baseline --no-ref-cuda $2 > gpurun_out/$3.json 2> gpurun_out/$3.err; echo "$3 rc=$?"; python tools/show_bench.py gpurun_out/$3.json 2>/dev/null | head -1 | cut -c1-140; }
run 29532 "" r2_c3_2gpu_fused
run 29533 "--no-defer" r2_c3_2gpu_fused_nodefer
run 29534 "" r2_c3_2gpu_fused_b
