#!/bin/bash
# GPU job K: smoke, parity suite, default bench, warm launch list of the config-3 step
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/r2_smoke.txt
if [ $rc -ne 0 ]; then exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2_pytest5.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest5.txt
timeout 500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_all_k.json 2> gpurun_out/r2_bench_all_k.err; echo "bench all rc=$?"; tail -c 300 gpurun_out/r2_bench_all_k.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv --log-file gpurun_out/r2_launches_c3_warm_k.csv python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --graph off > gpurun_out/r2_ncu_c3_k.log 2>&1; echo "ncu rc=$?"
