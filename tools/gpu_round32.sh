#!/bin/bash
mkdir -p gpurun_out
INFLIGHT=3 timeout 600 python tools/phase_times.py > gpurun_out/phase_times_if3.txt 2>&1; tail -21 gpurun_out/phase_times_if3.txt | head -17; tail -2 gpurun_out/phase_times_if3.txt
