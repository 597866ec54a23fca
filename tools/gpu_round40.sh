#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -p no:cacheprovider -rf --tb=short -x -k "tcgen05_attention" > gpurun_out/pytest_attn.log 2>&1; tail -25 gpurun_out/pytest_attn.log
timeout 300 python tools/attn_prefill_bench.py > gpurun_out/attn_prefill.txt 2>&1; tail -8 gpurun_out/attn_prefill.txt
