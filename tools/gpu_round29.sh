#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
run() { n=$1; lib=$2; shift 2
  DNET_B200_LIB=$lib timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --no-e2e "$@" > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err; echo "$n: $(grep -E 'value' gpurun_out/bench_$n.err | cut -c1-60) $(grep -o '"check": {[^}]*}' gpurun_out/bench_$n.json)"
}
D=$PWD/dnet_b200/lib/libdnet_b200.so
run base $PWD/dnet_b200/lib/ab/libdnet_b200_base.so
run mma $D
run null $D --mk-flags 4
run mma2 $D
CALIB=1 timeout 600 python tools/phase_times.py > gpurun_out/phase_times.txt 2>&1; tail -22 gpurun_out/phase_times.txt
