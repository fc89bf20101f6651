#!/bin/bash
# The encode part of tools/make_profiles.sh alone (C2 random bytes, C5 ACGTN): launch list + one --set full capture of
# the second step, summarised on the box.
TAG=${1:-r2c}
OUT=gpurun_out/profiles_$TAG
B=text_compression_b200/csrc/build
mkdir -p $OUT
summ() {
  local rep=$1 label=$2; shift 2
  python tools/ncu_summary.py $rep $label > $OUT/${TAG}_ncu_full_summary_${label}.csv
  : > $OUT/${TAG}_ncu_stalls_${label}.txt
  for K in "$@"; do
    python tools/ncu_stalls.py $rep $K 10 >> $OUT/${TAG}_ncu_stalls_${label}.txt 2>/dev/null
  done
}
for KIND in bytes acgtn; do
  python tools/one_step.py $KIND 16777216 2 || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_${KIND}.csv \
      python tools/one_step.py $KIND 16777216 2 > /dev/null 2>&1
  NL=$(grep -c '"gpu__time_duration.sum"' $OUT/${TAG}_launches_${KIND}.csv)
  ncu --set full --clock-control none --import-source on --launch-skip $((NL / 2)) -o /tmp/prof_$KIND -f \
      python tools/one_step.py $KIND 16777216 2 > $OUT/${TAG}_ncu_full_${KIND}.log 2>&1
  summ /tmp/prof_$KIND.ncu-rep $KIND part_kernel final_sort mtf3_replay mtfa_replay
done
python bench.py --steps 3 --warmup 3 --fm 0 --locate 0 --c1 0 --decode 0 > $OUT/${TAG}_bench_short.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --fm 0 --locate 0 --c1 0 --decode 0 > /dev/null 2>&1
ls $OUT
