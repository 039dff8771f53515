#!/bin/bash
# GPU job U: warm / cold ncu launch lists of the config-3 step (per-launch device times)
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv --log-file gpurun_out/r2_launches_c3_warm_u.csv python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --graph off > gpurun_out/r2_ncu_c3_u.log 2>&1; echo "ncu warm rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c3_cold_u.csv python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --graph off > gpurun_out/r2_ncu_c3_u2.log 2>&1; echo "ncu cold rc=$?"
python tools/step_launches.py gpurun_out/r2_launches_c3_cold_u.csv gpurun_out/r2_launches_c3_warm_u.csv
