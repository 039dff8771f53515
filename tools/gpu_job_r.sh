#!/bin/bash
# GPU job R: weight-side kernels on an auxiliary stream: parity suite, config 3 / 4 with and without it, default bench
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/r2_smoke.txt
if [ $rc -ne 0 ]; then exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2_pytest_r.txt 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_r.txt
run() { timeout 300 python bench.py --config $1 --steps 100 --warmup 5 --no-cpu-baseline --no-ref-cuda $2 > gpurun_out/$3.json 2> gpurun_out/$3.err; echo "$3 rc=$?"; python tools/show_bench.py gpurun_out/$3.json 2>/dev/null | head -1 | cut -c1-140; }
run c3 "" r2_c3_aux
run c3 "--plan-overrides 32" r2_c3_noaux
run c3 "" r2_c3_aux_b
run c3 "--plan-overrides 32" r2_c3_noaux_b
run c4 "" r2_c4_aux
run c4 "--plan-overrides 32" r2_c4_noaux
