#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hop.py -m gpu -q --timeout 300 -p no:cacheprovider -rf --tb=short > gpurun_out/pytest_hop.log 2>&1; tail -15 gpurun_out/pytest_hop.log
for f in 1 0; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$f bench.py --gpus 2 --steps 128 --warmup 8 --fused-hop $f > gpurun_out/bench_n2_f$f.json 2> gpurun_out/bench_n2_f$f.err
echo "n2 fused=$f exit $?"; grep -E "bench\]|Error|error" gpurun_out/bench_n2_f$f.err | tail -5; python -c "
import json; d=json.load(open('gpurun_out/bench_n2_f$f.json')); s=d['roofline']['single_sequence']; print(d['value'], d['ms_per_step'], 'single', s['tok_s'], 'hop_us', s['ring_hop_us'], 'e2e', (d['e2e'] or {}).get('value'), d['check'], d['config']['hop'])"
done
timeout 600 python bench.py --steps 128 --warmup 8 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; grep -E "value|e2e|cpu_base" gpurun_out/bench_n1.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['check'], d['roofline']['frac'])"
