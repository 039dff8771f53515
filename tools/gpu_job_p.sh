#!/bin/bash
# GPU job P (1 GPU): does the dX-first / dW-last order of the backward change the single-GPU step?
mkdir -p gpurun_out
run() { timeout 300 python bench.py --config c3 --steps 100 --warmup 5 --no-cpu-baseline --no-ref-cuda $1 > gpurun_out/$2.json 2> gpurun_out/$2.err; echo "$2 rc=$?"; python tools/show_bench.py gpurun_out/$2.json 2>/dev/null | head -1 | cut -c1-140; }
run "" r2_c3_1gpu_order_bwd
run "--force-defer" r2_c3_1gpu_order_defer
run "" r2_c3_1gpu_order_bwd_b
run "--force-defer" r2_c3_1gpu_order_defer_b
