#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/prefill_bench.py > gpurun_out/prefill.txt 2>&1; tail -6 gpurun_out/prefill.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_gemm_tc' -s 40 -c 4 -o gpurun_out/prof_gemm env LAYERS=4 python tools/prefill_bench.py > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
