#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/phase_times.py > gpurun_out/phase_times_park.txt 2>&1; tail -21 gpurun_out/phase_times_park.txt | head -17; tail -2 gpurun_out/phase_times_park.txt
# full default bench line (value + e2e + cpu baseline)
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -4 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json | head -c 1500; echo
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | head -c 600; echo
# launch list of the default command (short) and one full capture of the step kernel
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'dn::' -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_shard_step' -s 3 -c 1 -o gpurun_out/prof_step python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log
