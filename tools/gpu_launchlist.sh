#!/bin/bash
# per-kernel device time of one bench step (split pipeline unless SCPP_SLICE says otherwise)
mkdir -p gpurun_out
SCPP_SLICE=${SCPP_SLICE:--1} timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_split.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/launches_split.log 2>&1; tail -1 gpurun_out/launches_split.log | cut -c1-150
