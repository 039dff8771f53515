#!/bin/bash
# GPU job M: smoke, parity suite, default bench (dX epilogue with two inputs in flight, lagged e2e loss read)
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/r2_smoke.txt
if [ $rc -ne 0 ]; then exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2_pytest_m.txt 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_m.txt
timeout 500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_all_m.json 2> gpurun_out/r2_bench_all_m.err; echo "bench all rc=$?"; tail -c 300 gpurun_out/r2_bench_all_m.err
python tools/show_bench.py gpurun_out/r2_bench_all_m.json 2>/dev/null | head -40
