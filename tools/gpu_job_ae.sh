#!/bin/bash
# GPU job AE (2 GPUs): config 2 and config 4 under data parallelism at the end of the round (sanity of the N > 1 paths)
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --config $2 --steps $3 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/$4.json 2> gpurun_out/$4.err; echo "$4 rc=$?"; python tools/show_bench.py gpurun_out/$4.json 2>/dev/null | head -1 | cut -c1-150; }
run 29571 c2 10 r2_c2_2gpu_final
run 29572 c4 100 r2_c4_2gpu_final
