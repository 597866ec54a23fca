#!/bin/bash
mkdir -p gpurun_out
run() { n=$1; shift
  timeout 200 python bench.py --steps 128 --warmup 8 --no-cpu --no-e2e "$@" > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err; echo "$n rc=$?: $(grep -E 'value' gpurun_out/bench_$n.err | cut -c1-60) $(grep -o 'calibrated in.*' gpurun_out/bench_$n.err | cut -c1-200)"
}
run base
run hi4 --inflight-hi 4
run hi5 --inflight-hi 5
run hi6 --inflight-hi 6
run cal --calibrate 1
run cal_hi5 --calibrate 1 --inflight-hi 5
run base2
