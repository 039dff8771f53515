#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/r2_smoke.txt
if [ $rc -ne 0 ]; then exit 1; fi
C3="python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --graph off"
timeout 200 $C3 > gpurun_out/r2_plain_c3.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c3.csv $C3 > gpurun_out/r2_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
C2="python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda"
timeout 200 $C2 > gpurun_out/r2_plain_c2.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_c2.csv $C2 > gpurun_out/r2_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
C4="python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --graph off"
timeout 200 $C4 > gpurun_out/r2_plain_c4.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c4.csv $C4 > gpurun_out/r2_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 300 python bench.py --config c3 --steps 50 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_c3_d.json 2> gpurun_out/r2_c3_d.err; echo "c3 rc=$?"
timeout 300 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_c2_d.json 2> gpurun_out/r2_c2_d.err; echo "c2 rc=$?"
