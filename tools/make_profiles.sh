#!/bin/bash
# Runs on the GPU box (gpurun): plain run first, then the ncu launch list and one --set full
# capture of the second C2 step (random bytes) and of a C5-style step (ACGTN).  The reports are
# summarised here (they are too large to bring back) and only text lands in gpurun_out/.
TAG=${1:-r1}
OUT=gpurun_out/profiles_$TAG
mkdir -p $OUT
for KIND in bytes acgtn; do
  python tools/one_step.py $KIND 16777216 2 || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_${KIND}.csv \
      python tools/one_step.py $KIND 16777216 2 > /dev/null 2>&1
  NL=$(grep -c '"gpu__time_duration.sum"' $OUT/${TAG}_launches_${KIND}.csv)
  ncu --set full --clock-control none --import-source on --launch-skip $((NL / 2)) -o /tmp/prof_$KIND -f \
      python tools/one_step.py $KIND 16777216 2 > $OUT/${TAG}_ncu_full_${KIND}.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$KIND.ncu-rep $KIND > $OUT/${TAG}_ncu_full_summary_${KIND}.csv
  : > $OUT/${TAG}_ncu_stalls_${KIND}.txt
  : > $OUT/${TAG}_ncu_lines_${KIND}.txt
  for K in mtf2_replay final_sort part_kernel uk_keys rle_emit seg_hist mtf2_lastocc mtfs_replay mtfs_summary sa_pack; do
    O=text_compression_b200/csrc/build/sufsort.o
    case $K in mtf*) O=text_compression_b200/csrc/build/mtf.o;; rle*) O=text_compression_b200/csrc/build/rle.o;; esac
    python tools/ncu_stalls.py /tmp/prof_$KIND.ncu-rep $K 10 >> $OUT/${TAG}_ncu_stalls_${KIND}.txt 2>/dev/null
    python tools/ncu_lines.py /tmp/prof_$KIND.ncu-rep $K $O 14 >> $OUT/${TAG}_ncu_lines_${KIND}.txt 2>/dev/null
  done
done
# SASS evidence: mnemonics of the hot kernels (no tensor-core / TMA instructions on this path)
cuobjdump -sass text_compression_b200/libtc_b200.so | grep -E "^\s+/\*[0-9a-f]{4}\*/" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -40 > $OUT/${TAG}_sass_mnemonics.txt
ls -la $OUT
echo profiles done
