#!/bin/bash
# GPU job W: gather forward with prefetch depth 2 at 64 accumulators (fewer spills): parity at the sweep geometry + the config-5 point
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sweep_geometry or fc_vs_oracle or full_size" > gpurun_out/r2_pytest_w.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_w.txt
for S in 64 32; do timeout 600 python bench.py --config c5 --n 4 --segments $S --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda 2>/dev/null | tail -1 > gpurun_out/r2_c5_w_$S.json; python tools/show_bench.py gpurun_out/r2_c5_w_$S.json | head -2 | cut -c1-260; done
timeout 600 python bench.py --config c5 --n 8 --segments 64 --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda 2>/dev/null | tail -1 > gpurun_out/r2_c5_w_8_64.json; python tools/show_bench.py gpurun_out/r2_c5_w_8_64.json | head -2 | cut -c1-260
