#!/bin/bash
# GPU job AF: end-of-round sanity on the committed sources: smoke + the config-3 line
# (the run before it tried a grid-stride prep kernel -- one wave of CTAs walking the tiles: slower, 39 -> 44 us at
#  config 3, 0.315 -> 0.375 ms at config 2 -- reverted)
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/r2_smoke.txt
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m pytest tests -m gpu -q -x -k "golden or oracle or deferred or sink" > gpurun_out/r2_pytest_af.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_af.txt
timeout 300 python bench.py --config c3 --steps 100 --warmup 5 --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_c3_af.json 2>/dev/null; python tools/show_bench.py gpurun_out/r2_c3_af.json | head -1 | cut -c1-150
