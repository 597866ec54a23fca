#!/bin/bash
# One gpurun call that re-validates round 2 on ONE GPU: GPU tests, smoke, the default bench line, the reference arm,
# the ncu launch list of the bench command and one full ncu capture of the step kernel (summaries -> profiles/).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -rf --tb=short -x > gpurun_out/r02_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -3 gpurun_out/r02_bench_n1.err | cut -c1-600
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; head -c 500 gpurun_out/r02_bench_reference_arm.json; echo
if [ "${NCU:-1}" = "1" ]; then
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'k_' -c 400 --csv --log-file gpurun_out/r02_launches_default_bench_raw.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-single > gpurun_out/r02_ncu_launch.log 2>&1; tail -2 gpurun_out/r02_ncu_launch.log | cut -c1-300
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_shard_step' -s 4 -c 1 -o gpurun_out/r02_prof_step python bench.py --steps 4 --warmup 3 --no-cpu --no-single > gpurun_out/r02_ncu_step.log 2>&1; tail -2 gpurun_out/r02_ncu_step.log | cut -c1-300
  ncu -i gpurun_out/r02_prof_step.ncu-rep --page raw --csv > gpurun_out/r02_ncu_step_raw.csv 2>/dev/null; wc -l gpurun_out/r02_ncu_step_raw.csv
fi
# driver settings (what the round-end bench runs) and the MoE config on one GPU
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_k20.json 2> gpurun_out/r02_bench_n1_k20.err; python -c "import json; d=json.load(open('gpurun_out/r02_bench_n1_k20.json')); print('K=20', d['value'], d['e2e']['value'], d['roofline']['frac'], d.get('check',{}).get('oracle_full_depth'))" | cut -c1-900
timeout 600 python bench.py --config moe --steps 32 --warmup 4 --in-flight 4 > gpurun_out/r02_bench_moe_n1.json 2> gpurun_out/r02_bench_moe_n1.err; tail -2 gpurun_out/r02_bench_moe_n1.err | cut -c1-400; head -c 700 gpurun_out/r02_bench_moe_n1.json; echo
