#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider -rf --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
run() { n=$1; lib=$2; shift 2
  DNET_B200_LIB=$lib timeout 250 python bench.py --steps 64 --warmup 8 --no-cpu --no-e2e "$@" > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err; echo "$n rc=$?: $(grep -E 'value' gpurun_out/bench_$n.err | cut -c1-60) $(grep -o '"check": {[^}]*}' gpurun_out/bench_$n.json)"
}
D=$PWD/dnet_b200/lib/libdnet_b200.so
B=$PWD/dnet_b200/lib/ab/libdnet_b200_prev.so
run ctx128_prev $B
run ctx128 $D
run ctx128_prev2 $B
run ctx128b $D
run ctx512 $D --prompt-len 512
run ctx2k $D --prompt-len 2048
run ctx8k $D --prompt-len 8192
run ctx8k_c64 $D --prompt-len 8192 --attn-chunk 64
