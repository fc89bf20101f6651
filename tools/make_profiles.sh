#!/bin/bash
# Runs on the GPU box (gpurun): every capture is preceded by the same command without ncu.  The reports are
# summarised here (they are too large to bring back) and only text lands in gpurun_out/profiles_$TAG/.
#   encode steps (C2 random bytes, C5 ACGTN): launch list + one --set full capture of the second step
#   the LSD + prefix-doubling suffix sort (TC_B200_NO_MSD=1), the inverse chain (decode), the FM-index kernels
TAG=${1:-r2}
OUT=gpurun_out/profiles_$TAG
B=text_compression_b200/csrc/build
mkdir -p $OUT
summ() {  # report label kernels...
  local rep=$1 label=$2; shift 2
  python tools/ncu_summary.py $rep $label > $OUT/${TAG}_ncu_full_summary_${label}.csv
  : > $OUT/${TAG}_ncu_stalls_${label}.txt
  : > $OUT/${TAG}_ncu_lines_${label}.txt
  for K in "$@"; do
    O=$B/sufsort.o
    case $K in mtf*) O=$B/mtf.o;; rle*) O=$B/rle.o;; fm_*) O=$B/fm.o;; inv_*|cs_*|bwt_*) O=$B/bwt.o;; rs_*) O=$B/radix.o;; esac
    python tools/ncu_stalls.py $rep $K 10 >> $OUT/${TAG}_ncu_stalls_${label}.txt 2>/dev/null
    python tools/ncu_lines.py $rep $K $O 14 >> $OUT/${TAG}_ncu_lines_${label}.txt 2>/dev/null
  done
}
for KIND in bytes acgtn; do
  python tools/one_step.py $KIND 16777216 2 || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_${KIND}.csv \
      python tools/one_step.py $KIND 16777216 2 > /dev/null 2>&1
  NL=$(grep -c '"gpu__time_duration.sum"' $OUT/${TAG}_launches_${KIND}.csv)
  ncu --set full --clock-control none --import-source on --launch-skip $((NL / 2)) -o /tmp/prof_$KIND -f \
      python tools/one_step.py $KIND 16777216 2 > $OUT/${TAG}_ncu_full_${KIND}.log 2>&1
  summ /tmp/prof_$KIND.ncu-rep $KIND mtf3_replay mtf3_starts mtf3_tile_last mtfa_replay mtfa_summary final_sort part_kernel uk_keys \
       rle_emit_tiled seg_hist sa_pack
done
# LSD + prefix doubling path (the path of the 100 Mbp / 1 Gbp index builds and of correlated text), 16 MiB ACGTN
TC_B200_NO_MSD=1 python tools/one_step.py acgtn 16777216 1 || exit 1
TC_B200_NO_MSD=1 ncu --set full --clock-control none --import-source on -k regex:"rs_|sa_" -c 40 -o /tmp/prof_lsd -f \
    python tools/one_step.py acgtn 16777216 1 > $OUT/${TAG}_ncu_full_lsd.log 2>&1
summ /tmp/prof_lsd.ncu-rep lsd rs_scatter rs_hist sa_keys2 sa_update
# inverse chain
python tools/decode_times.py bytes > $OUT/${TAG}_decode_times_bytes.txt || exit 1
python tools/decode_times.py acgtn > $OUT/${TAG}_decode_times_acgtn.txt || exit 1
ncu --set full --clock-control none --import-source on -k regex:"inv_|cs_|mtfd|rle_expand|rle_len" -c 40 -o /tmp/prof_dec -f \
    python tools/decode_times.py bytes > $OUT/${TAG}_ncu_full_decode.log 2>&1
summ /tmp/prof_dec.ncu-rep decode inv_walk1 inv_jump_all inv_walk2 cs_scatter mtfd3_perm mtfd3_replay mtfd_tile_chain rle_expand
# FM-index: 100 Mbp (index mostly L2-resident) and 1 Gbp (1.98 GB image: HBM)
python tools/fm_step.py 100000000 1000000 200000 > $OUT/${TAG}_fm_step_100M.txt || exit 1
ncu --set full --clock-control none --import-source on -k regex:"fm_" -c 12 -o /tmp/prof_fm -f \
    python tools/fm_step.py 100000000 1000000 200000 > $OUT/${TAG}_ncu_full_fm100M.log 2>&1
summ /tmp/prof_fm.ncu-rep fm100M fm_count fm_locate fm_planes
python tools/fm_step.py 1000000000 1000000 200000 > $OUT/${TAG}_fm_step_1G.txt || exit 1
ncu --set full --clock-control none --import-source on -k regex:"fm_count|fm_locate" -c 4 -o /tmp/prof_fm1g -f \
    python tools/fm_step.py 1000000000 1000000 200000 > $OUT/${TAG}_ncu_full_fm1G.log 2>&1
summ /tmp/prof_fm1g.ncu-rep fm1G fm_count fm_locate
# what serves the sectors: DRAM vs L2 bytes per launch
ncu -i /tmp/prof_fm.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum > $OUT/${TAG}_fm_traffic_fm100M.csv 2>/dev/null
ncu -i /tmp/prof_fm1g.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum > $OUT/${TAG}_fm_traffic_fm1G.csv 2>/dev/null
# the bench command itself: launch list (shares of the step, cold and serialised)
python bench.py --steps 3 --warmup 3 --fm 0 --locate 0 --c1 0 --decode 0 > $OUT/${TAG}_bench_short.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --fm 0 --locate 0 --c1 0 --decode 0 > /dev/null 2>&1
ls -la $OUT
echo profiles done
