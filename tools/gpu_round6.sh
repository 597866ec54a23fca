#!/bin/bash
mkdir -p gpurun_out
for f in 0 1 2 3; do
timeout 600 python bench.py --steps 128 --warmup 8 --megakernel 1 --no-cpu --no-e2e --pf-depth 0 --mk-flags $f > gpurun_out/bench_f$f.json 2> gpurun_out/bench_f$f.err; echo "flags $f exit $?"; grep -E "value" gpurun_out/bench_f$f.err
done
