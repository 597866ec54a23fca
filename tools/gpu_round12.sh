#!/bin/bash
mkdir -p gpurun_out
for d in 8 16 24 48; do
timeout 600 python bench.py --steps 128 --warmup 8 --megakernel 1 --no-cpu --no-e2e --pf-depth $d > gpurun_out/bench_pf$d.json 2> gpurun_out/bench_pf$d.err; echo "pf $d exit $?"; grep -E "value" gpurun_out/bench_pf$d.err
done
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_shard_step' -s 2 -c 1 -o gpurun_out/prof_step python bench.py --steps 2 --warmup 3 --prompt-len 4 --no-e2e --no-cpu --megakernel 1 --pf-depth 16 > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
