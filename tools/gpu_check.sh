#!/bin/bash
# One gpurun call that re-validates the round: GPU tests, smoke, the default bench line, the
# reference arm, the ncu launch list of the bench command and one full ncu capture of the step kernel.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider -rf --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -4 gpurun_out/bench_n1.err | cut -c1-400
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; head -c 400 gpurun_out/bench_ref.json; echo
if [ "${NCU:-1}" = "1" ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'k_' -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_shard_step' -s 3 -c 1 -o gpurun_out/prof_step python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log
fi
