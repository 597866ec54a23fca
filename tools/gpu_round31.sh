#!/bin/bash
mkdir -p gpurun_out
run() { n=$1; shift
  timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --no-e2e --calibrate 0 "$@" > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err; echo "$n: $(grep -E 'value' gpurun_out/bench_$n.err | cut -c1-60)"
}
run if0
run if2 --inflight 2
run if3 --inflight 3
run if4 --inflight 4
run if5 --inflight 5
run if0b
