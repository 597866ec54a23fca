#!/bin/bash
mkdir -p gpurun_out
DNET_B200_LIB=$PWD/dnet_b200/lib/ab/libdnet_b200_prev.so INFLIGHT=3 timeout 200 python tools/phase_times.py > gpurun_out/pt_prev.txt 2>&1; sed -n 2,17p gpurun_out/pt_prev.txt
INFLIGHT=3 timeout 200 python tools/phase_times.py > gpurun_out/pt_new.txt 2>&1; sed -n 2,17p gpurun_out/pt_new.txt
