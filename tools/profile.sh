#!/bin/bash
# ncu evidence for one round (run under gpurun, ONE GPU): launch list of the bench step + full captures of the
# dominant kernels.  usage: bash tools/profile.sh <tag>   -> gpurun_out/<tag>_*.{csv,ncu-rep}
tag=${1:-r1c}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fc_tc2 -s 12 -c 2 -f -o gpurun_out/${tag}_fc $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ls_ -s 3 -c 1 -f -o gpurun_out/${tag}_ls $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ofdm_r16 -s 3 -c 1 -f -o gpurun_out/${tag}_ofdm python tools/bench_ofdm.py > /dev/null 2>&1
ls -la gpurun_out/
