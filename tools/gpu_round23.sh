#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/offload_bench.py > gpurun_out/offload.json 2> gpurun_out/offload.err; echo "exit $?"; tail -5 gpurun_out/offload.err; cat gpurun_out/offload.json
