#!/bin/bash
# GPU job AD (4 GPUs): config 3 under data parallelism (completes the 1 / 2 / 4 / 8 table)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --config c3 --steps 100 --warmup 5 --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_c3_4gpu_final.json 2> gpurun_out/r2_c3_4gpu_final.err; echo "rc=$?"; python tools/show_bench.py gpurun_out/r2_c3_4gpu_final.json 2>/dev/null | head -1 | cut -c1-150
