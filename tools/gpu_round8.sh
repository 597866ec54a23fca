#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/phase_times.py > gpurun_out/phase_times.txt 2>&1; cat gpurun_out/phase_times.txt | tail -20
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_shard_step' -s 2 -c 1 -o gpurun_out/prof_step python bench.py --steps 2 --warmup 3 --prompt-len 4 --no-e2e --no-cpu --megakernel 1 --pf-depth 0 > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
