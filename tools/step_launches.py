#!/usr/bin/env python
"""Per-launch device times of ONE config-3 step out of an ncu launch list (gpu__time_duration.sum, --csv):
the launches between the last three prep kernels before the last loss kernel and the first layer's dW fix-up.
Usage: step_launches.py cold.csv [warm.csv]   (two files: printed side by side)"""
import csv
import sys


def step(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = []
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        seq.append((r[ki].split("(")[0][:58], v / 1000 if r[ui].startswith("ns") else v))
    last = [i for i, (n, _) in enumerate(seq) if "nll_loss_forward" in n][-1]
    start = [i for i, (n, _) in enumerate(seq[:last]) if "pw_prep_fc" in n][-3]
    out, fixes = [], 0
    for n, v in seq[start:]:
        out.append((n, v))
        if "pw_tc_fix_dw" in n:
            fixes += 1
            if fixes == 2:
                break
    return out


cols = [step(p) for p in sys.argv[1:]]
for k, (name, v) in enumerate(cols[0]):
    extra = "".join("%9.2f" % c[k][1] if k < len(c) and c[k][0] == name else "        ?" for c in cols[1:])
    print("%9.2f%s  %s" % (v, extra, name))
print("".join("%9.2f" % sum(v for _, v in c) for c in cols), " total (us)")
