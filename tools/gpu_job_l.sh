#!/bin/bash
# GPU job L: ncu --set full of the small latency-bound kernels of the config-3 step
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --cache-control none -k regex:"pw_fwd_kernel|pw_prep_fc_kernel|pw_dx_kernel|pw_dw_dense_kernel|pw_tc_finalize_kernel" -c 16 -f -o gpurun_out/r2_prof_small python bench.py --config c3 --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda --graph off > gpurun_out/r2_ncu_small.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_small.log
