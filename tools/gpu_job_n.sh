#!/bin/bash
# GPU job N (2 GPUs): config 3 under data parallelism with and without the largest-message-first deferral of dW
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --config c3 --steps 100 --warmup 5 --no-cpu-baseline --no-ref-cuda $2 > gpurun_out/$3.json 2> gpurun_out/$3.err; echo "$3 rc=$?"; tail -c 200 gpurun_out/$3.err; python tools/show_bench.py gpurun_out/$3.json | head -2; }
run 29511 "" r2_c3_2gpu_defer
run 29512 "--no-defer" r2_c3_2gpu_nodefer
run 29513 "" r2_c3_2gpu_defer_b
run 29514 "--no-defer" r2_c3_2gpu_nodefer_b
timeout 300 python bench.py --config c3 --steps 100 --warmup 5 --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_c3_1gpu_n.json 2>/dev/null; python tools/show_bench.py gpurun_out/r2_c3_1gpu_n.json | head -1
