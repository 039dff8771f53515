#!/bin/bash
# GPU job T: e2e leg with three input buffers
mkdir -p gpurun_out
run() { timeout 300 python bench.py --config $1 --steps $2 --warmup 5 --no-cpu-baseline --no-ref-cuda > gpurun_out/$3.json 2> gpurun_out/$3.err; echo "$3 rc=$?"; tail -c 300 gpurun_out/$3.err; python tools/show_bench.py gpurun_out/$3.json 2>/dev/null | head -1 | cut -c1-140; python -c "
import json,sys; d=json.loads(open('gpurun_out/$3.json').read().strip().splitlines()[-1]); e=d['e2e']; print('   e2e ms', round(e['ms_per_step'],4), 'h2d alone ms', round(e['h2d_ms_alone'],4), 'GB/s', round(e['h2d_gbs_alone'],1))"; }
run c3 20 r2_c3_e2e3
run c3 100 r2_c3_e2e3_b
run c4 100 r2_c4_e2e3
run c2 10 r2_c2_e2e3
