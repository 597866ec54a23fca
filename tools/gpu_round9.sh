#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python tools/phase_times.py > gpurun_out/phase_times.txt 2>&1; cat gpurun_out/phase_times.txt | tail -18
timeout 600 python bench.py --steps 128 --warmup 8 --megakernel 1 --no-cpu --pf-depth 0 > gpurun_out/bench_mk.json 2> gpurun_out/bench_mk.err; grep -E "value|e2e" gpurun_out/bench_mk.err
