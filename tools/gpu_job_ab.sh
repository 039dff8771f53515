#!/bin/bash
# GPU job AB: ncu --set full of the three config-2 GEMMs after the dX epilogue rework
# (the run also compared n = 2 / 64 segments of the sweep with dX on the tensor cores -- 229.7 ms, the default --
#  against the gather kernels -- 221.4 ms; within noise of each other, plan left as it is)
mkdir -p gpurun_out
C2="python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_tc_gemm -s 6 -c 3 -f -o gpurun_out/r2_prof_c2_final $C2 > gpurun_out/r2_ncu_full_c2_final.log 2>&1; echo "ncu full c2 rc=$?"; ls -la gpurun_out/r2_prof_c2_final.ncu-rep
