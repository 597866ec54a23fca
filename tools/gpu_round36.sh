#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider -rf --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
run() { n=$1; shift
  timeout 250 python bench.py --steps 64 --warmup 8 --no-cpu --no-e2e "$@" > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err; echo "$n rc=$?: $(grep -E 'value' gpurun_out/bench_$n.err | cut -c1-60)"
}
run ctx128
run ctx512 --prompt-len 512
run ctx2k --prompt-len 2048
run ctx8k --prompt-len 8192
run ctx8k_c64 --prompt-len 8192 --attn-chunk 64
run ctx8k_c128 --prompt-len 8192 --attn-chunk 128
run ctx2k_c64 --prompt-len 2048 --attn-chunk 64
