#!/bin/bash
# ncu evidence for the LMMSE smoother: launch list (both bench cases) + a full capture of mid-factorisation panel kernels
tag=${1:-r1c}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_lmmse_launches.csv python tools/bench_lmmse.py > gpurun_out/${tag}_lmmse_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lmmse_panel -s 4 -c 2 -f -o gpurun_out/${tag}_lmmse_panel python tools/bench_lmmse.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lmmse_backsub -s 0 -c 1 -f -o gpurun_out/${tag}_lmmse_backsub python tools/bench_lmmse.py > /dev/null 2>&1
ls -la gpurun_out | grep lmmse
