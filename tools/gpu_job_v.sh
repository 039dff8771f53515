#!/bin/bash
# GPU job V: ncu --set full of the gather forward / dX / dW kernels at the config-5 point n=4, S=64 (batch 8192)
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"pw_fwd_kernel|pw_dx_kernel|pw_dw_kernel" -c 3 -f -o gpurun_out/r2_prof_c5 python bench.py --config c5 --n 4 --segments 64 --batch 8192 --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_ncu_c5.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_c5.log; ls -la gpurun_out/r2_prof_c5.ncu-rep
