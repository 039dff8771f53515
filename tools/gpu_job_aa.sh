#!/bin/bash
# GPU job AA: ncu --set full of the expand + tcgen05 GEMM launches of one config-3 step (after the dX epilogue rework)
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --cache-control none -k regex:"pw_tc_gemm_kernel|pw_tc_expand_kernel" --launch-skip 21 -c 7 -f -o gpurun_out/r2_prof_c3_final python bench.py --config c3 --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda --graph off > gpurun_out/r2_ncu_c3_final.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/r2_prof_c3_final.ncu-rep
