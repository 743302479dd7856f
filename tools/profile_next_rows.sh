#!/bin/bash
# ncu evidence for the "next" rows (run under gpurun, ONE GPU): persistent OFDM kernel, LMMSE Toeplitz route
tag=${1:-r1e}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ofdm_r16_tma -s 17 -c 1 -f -o gpurun_out/${tag}_ofdm python tools/bench_ofdm.py > /dev/null 2>&1
MAMIMO_LMMSE_STREAMS=1 ncu --set full --clock-control none --import-source on -k regex:lmmse_s -s 0 -c 2 -f -o gpurun_out/${tag}_lmmse python tools/bench_lmmse.py > /dev/null 2>&1
MAMIMO_LMMSE_STREAMS=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lmmse -c 12 --csv --log-file gpurun_out/${tag}_lmmse_launches.csv python tools/bench_lmmse.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
ls -la gpurun_out | grep ${tag}
