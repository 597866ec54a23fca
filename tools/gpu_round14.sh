#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt; nvidia-smi topo -m >> gpurun_out/gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 8 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "n2 exit $?"; tail -25 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_shard_step' -s 2 -c 1 -o gpurun_out/prof_step python bench.py --steps 2 --warmup 3 --prompt-len 4 --no-e2e --no-cpu --megakernel 1 > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
