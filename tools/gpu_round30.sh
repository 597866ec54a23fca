#!/bin/bash
mkdir -p gpurun_out
run() { n=$1; shift
  timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --no-e2e "$@" > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err; echo "$n: $(grep -E 'value' gpurun_out/bench_$n.err | cut -c1-60)"
}
run cal1
run cal0 --calibrate 0
run nobar --mk-flags 8
run nobar_null --mk-flags 12
run nobar_cal0 --mk-flags 8 --calibrate 0
run cal1b
