#!/bin/bash
mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 128 --warmup 8 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "n$N exit $?"; grep -E "bench\]|rror" gpurun_out/bench_n$N.err | tail -12 | cut -c1-200; python -c "
import json
lines=[l for l in open('gpurun_out/bench_n$N.json') if l.strip().startswith('{')]
print('stdout lines', len(open('gpurun_out/bench_n$N.json').readlines()))
d=json.loads(lines[-1]); s=d['roofline']['single_sequence']; print(round(d['value'],1), round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'single', round(s['tok_s'],1), 'hop_us', round(s['ring_hop_us'],2), 'comp', [round(x,3) for x in s['per_rank_compute_ms']], 'e2e', round((d['e2e'] or {}).get('value',0),1), d['check']['token'], d['hop_timeout'], d['config']['hop'])"
