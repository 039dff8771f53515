#!/bin/bash
# GPU job A: parity tests, default bench line, launch lists (ncu, cold/serialised: shares only)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2_pytest2.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest2.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_all.json 2> gpurun_out/r2_bench_all.err; echo "bench all rc=$?"
C3="python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --graph off"
timeout 200 $C3 > gpurun_out/r2_plain_c3.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c3.csv $C3 > gpurun_out/r2_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
C2="python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda"
timeout 200 $C2 > gpurun_out/r2_plain_c2.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_c2.csv $C2 > gpurun_out/r2_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
