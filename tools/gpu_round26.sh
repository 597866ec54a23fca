#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for c in 32 64 128; do
timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --no-e2e --attn-chunk $c > gpurun_out/bench_ac$c.json 2> gpurun_out/bench_ac$c.err; echo "chunk $c"; grep -E "value" gpurun_out/bench_ac$c.err
done
CALIB=1 timeout 600 python tools/phase_times.py > gpurun_out/phase_times.txt 2>&1; tail -22 gpurun_out/phase_times.txt
