#!/bin/bash
mkdir -p gpurun_out
run() { n=$1; shift
  timeout 150 python bench.py --steps 128 --warmup 8 --no-cpu --no-e2e "$@" > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err; echo "$n rc=$?: $(grep -E 'value' gpurun_out/bench_$n.err | cut -c1-60) $(grep -o '"check": {[^}]*}' gpurun_out/bench_$n.json)"
}
run park0 --park 0
run park1 --park 1
timeout 500 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider -rf --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
run park1_if4 --park 1 --inflight 4
run park1_if0 --park 1 --inflight 0
run park1_if2 --park 1 --inflight 2
run park0b --park 0
