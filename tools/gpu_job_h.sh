#!/bin/bash
# ncu --set full captures of the tcgen05 GEMMs (one step's launches) at config 2 and config 3
mkdir -p gpurun_out
C2="python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda"
timeout 200 $C2 > gpurun_out/r2_plain_c2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_tc_gemm -s 6 -c 3 -f -o gpurun_out/r2_prof_c2 $C2 > gpurun_out/r2_ncu_full_c2.log 2>&1; echo "ncu full c2 rc=$?"
C3="python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --graph off"
timeout 200 $C3 > gpurun_out/r2_plain_c3.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_tc_gemm\|pw_tc_expand -s 21 -c 7 -f -o gpurun_out/r2_prof_c3 $C3 > gpurun_out/r2_ncu_full_c3.log 2>&1; echo "ncu full c3 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c3.csv $C3 > gpurun_out/r2_ncu_c3.log 2>&1; echo "ncu list c3 rc=$?"
ls -la gpurun_out/*.ncu-rep
