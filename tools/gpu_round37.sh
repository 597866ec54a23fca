#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 128 --warmup 8 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; tail -5 gpurun_out/bench_n2.err | cut -c1-400; head -c 1800 gpurun_out/bench_n2.json; echo
timeout 300 python -m pytest tests/test_gpu_hop.py -m gpu -q --timeout 120 -p no:cacheprovider -x 2>&1 | tail -3
