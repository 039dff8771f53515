#!/bin/bash
# GPU job O (2 GPUs): where do the +40 us of the 2-GPU config-3 step come from?  exchange skipped / NCCL / own kernels
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --config c3 --steps 100 --warmup 5 --no-cpu-baseline --no-ref-cuda $2 > gpurun_out/$3.json 2> gpurun_out/$3.err; echo "$3 rc=$?"; python tools/show_bench.py gpurun_out/$3.json 2>/dev/null | head -1 | cut -c1-140; }
run 29521 "--collective none" r2_c3_2gpu_none
run 29522 "--collective none --no-defer" r2_c3_2gpu_none_nodefer
run 29523 "--collective nccl" r2_c3_2gpu_nccl
run 29524 "" r2_c3_2gpu_p2p
