#!/bin/bash
# GPU job S: auxiliary stream only for large layers; e2e with the H2D copy timed alone
mkdir -p gpurun_out
run() { timeout 300 python bench.py --config $1 --steps 100 --warmup 5 --no-cpu-baseline --no-ref-cuda $2 > gpurun_out/$3.json 2> gpurun_out/$3.err; echo "$3 rc=$?"; python tools/show_bench.py gpurun_out/$3.json 2>/dev/null | head -1 | cut -c1-140; python -c "
import json,sys; d=json.loads(open('gpurun_out/$3.json').read().strip().splitlines()[-1]); e=d['e2e']; print('   e2e ms', round(e['ms_per_step'],4), 'h2d alone ms', round(e['h2d_ms_alone'],4), 'GB/s', round(e['h2d_gbs_alone'],1))"; }
run c3 "" r2_c3_aux1
run c3 "--plan-overrides 32" r2_c3_noaux1
run c3 "" r2_c3_aux1_b
run c3 "--plan-overrides 32" r2_c3_noaux1_b
timeout 300 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_c2_aux1.json 2>/dev/null; python tools/show_bench.py gpurun_out/r2_c2_aux1.json | head -2 | cut -c1-250
