#!/bin/bash
# GPU job Z (8 GPUs): config 3 under data parallelism, with and without the largest-message-first deferral; config 4
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --config $4 --steps 100 --warmup 5 --no-cpu-baseline --no-ref-cuda $2 > gpurun_out/$3.json 2> gpurun_out/$3.err; echo "$3 rc=$?"; python tools/show_bench.py gpurun_out/$3.json 2>/dev/null | head -1 | cut -c1-150; }
run 29541 "" r2_c3_8gpu_defer c3
run 29542 "--no-defer" r2_c3_8gpu_nodefer c3
run 29543 "" r2_c4_8gpu c4
