#!/bin/bash
# The inverse-chain part of tools/make_profiles.sh alone (after the decode kernels changed): run without ncu first,
# then one --set full capture, summarised on the box.
TAG=${1:-r2b}
OUT=gpurun_out/profiles_$TAG
B=text_compression_b200/csrc/build
mkdir -p $OUT
python tools/decode_times.py bytes > $OUT/${TAG}_decode_times_bytes.txt || exit 1
python tools/decode_times.py acgtn > $OUT/${TAG}_decode_times_acgtn.txt || exit 1
ncu --set full --clock-control none --import-source on -k regex:"inv_|cs_|mtfd|rle_expand|rle_len" -c 40 -o /tmp/prof_dec -f \
    python tools/decode_times.py bytes > $OUT/${TAG}_ncu_full_decode.log 2>&1
label=decode
rep=/tmp/prof_dec.ncu-rep
python tools/ncu_summary.py $rep $label > $OUT/${TAG}_ncu_full_summary_${label}.csv
: > $OUT/${TAG}_ncu_stalls_${label}.txt
: > $OUT/${TAG}_ncu_lines_${label}.txt
for K in inv_walk1 inv_jump_all inv_walk2 cs_scatter mtfd3_perm mtfd3_replay mtfd_tile_chain rle_expand; do
  O=$B/bwt.o
  case $K in mtf*) O=$B/mtf.o;; rle*) O=$B/rle.o;; esac
  python tools/ncu_stalls.py $rep $K 10 >> $OUT/${TAG}_ncu_stalls_${label}.txt 2>/dev/null
  python tools/ncu_lines.py $rep $K $O 14 >> $OUT/${TAG}_ncu_lines_${label}.txt 2>/dev/null
done
ls -la $OUT
