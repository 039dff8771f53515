#!/bin/bash
# BASELINE config 5: I=O=4096, B=32768, n in {2,4,8} x segments in {2,8,32,64}; one JSON line per point.
out=gpurun_out/sweep_c5.jsonl; : > $out
for n in 2 4 8; do for S in 2 8 32 64; do
  timeout 600 python bench.py --config c5 --n $n --segments $S --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda 2>gpurun_out/sweep_err_${n}_${S}.log | tail -1 >> $out || echo "{\"n\": $n, \"segments\": $S, \"error\": \"failed\"}" >> $out
done; done
python - <<'PY'
import json
for line in open('gpurun_out/sweep_c5.jsonl'):
    try: d=json.loads(line)
    except Exception: print('bad line', line[:100]); continue
    if 'error' in d and 'value' not in d: print(d); continue
    r=d.get('roofline') or {}; lay=r.get('layer',{}); km=r.get('kernel_ms',{})
    print(lay.get('n'), lay.get('segments'), round(d['ms_per_step'],1), 'ms', round(d['value']), 'samples/s', r.get('bound'), {k: round(v,1) for k,v in km.items() if v>0.05})
PY
