#!/bin/bash
# GPU job J: launch lists of the config-3 step, cold (ncu default: caches flushed before every kernel) and
# warm (--cache-control none: the caches are whatever the previous kernels of the step left)
mkdir -p gpurun_out
for cc in all none; do
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control $cc -c 600 --csv --log-file gpurun_out/r2_launches_c3_cc_$cc.csv python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --graph off > gpurun_out/r2_ncu_c3_j.log 2>&1; echo "ncu $cc rc=$?"
done
timeout 300 python bench.py --config c3 --steps 50 --warmup 5 --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_c3_j.json 2> gpurun_out/r2_c3_j.err; echo "bench c3 rc=$?"
