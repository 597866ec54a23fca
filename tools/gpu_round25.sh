#!/bin/bash
mkdir -p gpurun_out
CALIB=1 timeout 600 python tools/phase_times.py > gpurun_out/phase_times.txt 2>&1; tail -22 gpurun_out/phase_times.txt
timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --no-e2e > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; grep -E "value" gpurun_out/bench_a.err
