#!/bin/bash
mkdir -p gpurun_out
CALIB=1 timeout 300 python tools/phase_times.py > gpurun_out/phase_times_cal.txt 2>&1; tail -22 gpurun_out/phase_times_cal.txt
for c in 1 0; do
timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --calibrate $c > gpurun_out/bench_cal$c.json 2> gpurun_out/bench_cal$c.err; grep -E "value|e2e|calibrated" gpurun_out/bench_cal$c.err
done
