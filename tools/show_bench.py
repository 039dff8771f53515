#!/usr/bin/env python
"""Print the numbers of a bench.py JSON line (headline + sub-blocks) in a few lines."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])


def roof(r):
    if not r or "error" in r:
        return str(r)
    u = r.get("tensor_pipe_util", {}).get("frac")
    return (f"{r['bound']} {r['kernel'][:34]} frac {r['frac']:.4f}" + (f" util {u:.3f}" if u else "") +
            f" dom {r['dominant_pass']} " + str({k: round(v, 3) for k, v in r["kernel_ms"].items() if v > 0.004}))


print("headline", d["config"]["workload"][:36], "ms", round(d["ms_per_step"], 4), "value", round(d["value"]), "e2e",
      d.get("e2e") and round(d["e2e"]["value"]), "launches", d["gpu_launches_per_step"], "graph", d["config"]["cuda_graph"],
      "gpus", d["n_gpus"])
print("  ", roof(d.get("roofline")))
print("   clocks", d.get("clocks"))
if d.get("cpu_baseline"):
    print("   cpu", round(d["cpu_baseline"]["value"]), "ref cuda", (d.get("ref_torch_cuda") or {}).get("value"))
for k, v in (d.get("configs") or {}).items():
    if "error" in v:
        print(k, v)
        continue
    print(k, "ms", round(v["ms_per_step"], 3), "value", round(v["value"]), "e2e", v.get("e2e") and round(v["e2e"]["value"]),
          "launches", v["gpu_launches_per_step"], "graph", v["cuda_graph"])
    print("  ", roof(v.get("roofline")))
    print("   clocks", v["clocks"])
    if "ref_torch_cuda" in v:
        print("   ref cuda", v["ref_torch_cuda"].get("value"), v["ref_torch_cuda"].get("parity_rel_err"))
