#!/bin/bash
# One gpurun call: GPU tests, smoke, a short bench, the ncu launch list and one full capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 128 --warmup 8 > gpurun_out/bench_pdl1.json 2> gpurun_out/bench_pdl1.err; echo "bench exit $?"; tail -12 gpurun_out/bench_pdl1.err; cat gpurun_out/bench_pdl1.json
timeout 400 python bench.py --steps 128 --warmup 8 --pdl 0 --no-cpu --no-e2e > gpurun_out/bench_pdl0.json 2> gpurun_out/bench_pdl0.err; cat gpurun_out/bench_pdl0.json
timeout 400 python bench.py --steps 128 --warmup 8 --l2-prefetch-kb 0 --no-cpu --no-e2e > gpurun_out/bench_nopf.json 2> gpurun_out/bench_nopf.err; cat gpurun_out/bench_nopf.json
timeout 400 python bench.py --steps 128 --warmup 8 --l2-prefetch-kb 256 --no-cpu --no-e2e > gpurun_out/bench_pf256.json 2> gpurun_out/bench_pf256.err; cat gpurun_out/bench_pf256.json
# every launch of our kernels with its device time (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'dn::' -s 170 -c 340 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --prompt-len 4 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
# the top kernel once, full sections
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'OpGateUp' -s 40 -c 3 -o gpurun_out/prof_gateup python bench.py --steps 2 --warmup 3 --prompt-len 4 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
