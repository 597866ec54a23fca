#!/bin/bash
# Multi-GPU legs of round 2 (every command under its own timeout):  bash tools/gpu_multi_r02.sh <N>
#   N=2: config 4 (70B-dims layer swap) + decode at the driver's settings      N=4: long-prompt pipeline, MoE, decode
#   N=8: decode at the driver's settings + MoE (config 5 shape: 8 shards, 8 requests in flight)
N=${1:-2}
mkdir -p gpurun_out
run() {  # run <timeout> <port> <out> <bench args...>
  local to=$1 port=$2 out=$3; shift 3
  timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err
  echo "== $out rc=$?"; head -c 600 gpurun_out/$out.json; echo; tail -2 gpurun_out/$out.err | cut -c1-300
}
if [ "$N" = "2" ]; then
  run 600 29611 r02_bench_swap_n2 --config swap --steps 8 --warmup 3
  run 300 29621 r02_bench_n2 --steps 20 --warmup 5
fi
if [ "$N" = "4" ]; then
  run 400 29631 r02_bench_prefill_n4 --config prefill --prefill-len 32768 --steps 3 --warmup 1
  run 500 29641 r02_bench_moe_n4 --config moe --steps 32 --warmup 4 --in-flight 8
  run 300 29651 r02_bench_n4 --steps 20 --warmup 5
fi
if [ "$N" = "8" ]; then
  run 300 29661 r02_bench_n8 --steps 20 --warmup 5
  run 500 29671 r02_bench_moe_n8 --config moe --steps 32 --warmup 4 --in-flight 8
  run 300 29681 r02_bench_n8_k128 --steps 128 --warmup 8
fi
