#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2_pytest3.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest3.txt
timeout 500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_all_b.json 2> gpurun_out/r2_bench_all_b.err; echo "bench all rc=$?"; tail -c 800 gpurun_out/r2_bench_all_b.err
