"""Summarise an .ncu-rep on the CPU box (needs the `ncu` CLI, no GPU):

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep            # key metrics per captured launch
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep --source 1 # hottest SASS lines + mbarrier wait sites of launch #1

The second form needs a capture made with --import-source on / -lineinfo.  For the warp-specialised
tcgen05 GEMM the wait sites tell which role starves: spins on `full` = the MMA warp waits for data,
on `tmem_empty` = it waits for the epilogue, on `empty` = the producer waits for the MMA warp (good),
on `tmem_full` = the epilogue waits for the MMA warp (good)."""
import argparse
import csv
import io
import subprocess

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
    "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for k, r in enumerate(data):
        print(f"---- launch {k}: {r[idx['Kernel Name']][:100]}")
        for key in KEYS:
            if key in idx:
                print(f"  {key:72s} {r[idx[key]]} {units[idx[key]]}")


def source(rep, launch, top):
    rows = ncu_csv(rep, "source", ["--launch-skip", str(launch), "--launch-count", "1"])
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen, lines = set(), []
    for r in rows[2:]:
        if len(r) == len(hdr) and r[idx["Address"]] not in seen:
            seen.add(r[idx["Address"]])
            lines.append(r)
    num = lambda v: int(v) if v.isdigit() else 0
    S, I, SRC = idx["# Samples"], idx["Instructions Executed"], idx["Source"]
    total = sum(num(r[S]) for r in lines)
    print(rows[0][1][:100], "| samples", total, "| warp instructions", sum(num(r[I]) for r in lines))
    print("-- hottest lines")
    for r in sorted(lines, key=lambda r: -num(r[S]))[:top]:
        print(f"{num(r[S]):9d} {100.0 * num(r[S]) / max(total, 1):5.1f}%  x{num(r[I]):<11d} {r[SRC][:100]}")
    print("-- mbarrier wait sites (executions = spin iterations)")
    for r in lines:
        if "TRYWAIT" in r[SRC] and num(r[I]):
            print(f"{num(r[S]):9d} samples  x{num(r[I]):<11d} {r[SRC][:100]}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--source", type=int, default=None, help="index of the captured launch to analyse line by line")
    ap.add_argument("--top", type=int, default=30)
    a = ap.parse_args()
    if a.source is None:
        raw(a.report)
    else:
        source(a.report, a.source, a.top)
