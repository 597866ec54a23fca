#!/bin/bash
mkdir -p gpurun_out
run() { # name lib extra...
  n=$1; lib=$2; shift 2
  DNET_B200_LIB=$lib timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --no-e2e "$@" > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err; echo "$n: $(grep -E 'value' gpurun_out/bench_$n.err | cut -c1-60)"
}
D=$PWD/dnet_b200/lib/libdnet_b200.so
run a1 $D
run nopre $PWD/dnet_b200/lib/ab/libdnet_b200_nopre.so
run a2 $D
run null $D --mk-flags 4
run nopre2 $PWD/dnet_b200/lib/ab/libdnet_b200_nopre.so
run a3 $D
