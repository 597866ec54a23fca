#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider -rf --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
timeout 300 python tools/attn_prefill_bench.py > gpurun_out/attn_prefill.txt 2>&1; tail -7 gpurun_out/attn_prefill.txt | head -6
