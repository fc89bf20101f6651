#!/bin/bash
# Runs on the GPU box (gpurun): plain run first, then the ncu launch list and one --set full
# capture of a whole C2 step and of a C5-style (ACGTN) step.  Outputs land in gpurun_out/.
set -x
TAG=${1:-r1}
python tools/one_step.py bytes 16777216 2 || exit 1
python tools/one_step.py acgtn 16777216 2 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_bytes.csv \
    python tools/one_step.py bytes 16777216 2 > gpurun_out/ncu_launch_bytes.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_acgtn.csv \
    python tools/one_step.py acgtn 16777216 2 > gpurun_out/ncu_launch_acgtn.log 2>&1
# second step only (the first warms the arena): skip the first step's launches
NL=$(grep -c '"gpu__time_duration.sum"' gpurun_out/launches_${TAG}_bytes.csv)
ncu --set full --clock-control none --import-source on --launch-skip $((NL / 2)) -o gpurun_out/prof_${TAG}_bytes -f \
    python tools/one_step.py bytes 16777216 2 > gpurun_out/ncu_full_bytes.log 2>&1
NL=$(grep -c '"gpu__time_duration.sum"' gpurun_out/launches_${TAG}_acgtn.csv)
ncu --set full --clock-control none --import-source on --launch-skip $((NL / 2)) -o gpurun_out/prof_${TAG}_acgtn -f \
    python tools/one_step.py acgtn 16777216 2 > gpurun_out/ncu_full_acgtn.log 2>&1
echo profiles done
