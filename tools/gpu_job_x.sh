#!/bin/bash
# GPU job X: gather dX on the transposed gy: parity suite, config-5 points, config 3 / 4
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2_pytest_x.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_x.txt
for cfg in "4 64" "8 64" "8 32" "2 64"; do set -- $cfg; timeout 600 python bench.py --config c5 --n $1 --segments $2 --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda 2>/dev/null | tail -1 > gpurun_out/r2_c5_x_$1_$2.json; python tools/show_bench.py gpurun_out/r2_c5_x_$1_$2.json 2>/dev/null | head -2 | cut -c1-260; done
for c in c3 c4; do timeout 300 python bench.py --config $c --steps 100 --warmup 5 --no-cpu-baseline --no-ref-cuda 2>/dev/null | tail -1 > gpurun_out/r2_${c}_x.json; python tools/show_bench.py gpurun_out/r2_${c}_x.json 2>/dev/null | head -2 | cut -c1-260; done
