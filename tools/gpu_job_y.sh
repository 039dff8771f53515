#!/bin/bash
# GPU job Y: end-of-round record: smoke, parity suite, default bench line, reference arm, config-5 sweep
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/r2_smoke.txt
if [ $rc -ne 0 ]; then exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2_pytest_y.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_y.txt
timeout 500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_all_y.json 2> gpurun_out/r2_bench_all_y.err; echo "bench all rc=$?"; tail -c 300 gpurun_out/r2_bench_all_y.err
python tools/show_bench.py gpurun_out/r2_bench_all_y.json 2>/dev/null | cut -c1-250 | head -24
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_reference_y.json 2> gpurun_out/r2_bench_reference_y.err; echo "reference arm rc=$?"; cut -c1-400 gpurun_out/r2_bench_reference_y.json
bash tools/sweep_c5.sh
