#!/bin/bash
# GPU job AB: (1) n = 2 / 64 segments of the sweep with dX on the tensor cores (default) and on the gather kernels;
# (2) ncu --set full of the three config-2 GEMMs after the dX epilogue rework
mkdir -p gpurun_out
for ovr in 0 32; do timeout 300 python bench.py --config c5 --n 2 --segments 64 --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda --plan-overrides $ovr 2>/dev/null | tail -1 > gpurun_out/r2_c5_2_64_ovr$ovr.json; python tools/show_bench.py gpurun_out/r2_c5_2_64_ovr$ovr.json 2>/dev/null | head -2 | cut -c1-250; done
C2="python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_tc_gemm -s 6 -c 3 -f -o gpurun_out/r2_prof_c2_final $C2 > gpurun_out/r2_ncu_full_c2_final.log 2>&1; echo "ncu full c2 rc=$?"; ls -la gpurun_out/r2_prof_c2_final.ncu-rep
