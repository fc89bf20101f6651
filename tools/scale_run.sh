#!/bin/bash
# Scaling evidence on an N-GPU box (gpurun --gpus 8): pinned-copy ceiling and the C2 bench at 1, 2, 4, 8 GPUs.
# Output: gpurun_out/scale_r2/*.json
OUT=gpurun_out/scale_r2
mkdir -p $OUT
NG=${1:-8}
python tools/pcie_ceiling.py 2>/dev/null | grep '^{' > $OUT/r2_pcie_ceiling_1gpu.json
for n in 2 4 8; do
  [ $n -le $NG ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      tools/pcie_ceiling.py 2>/dev/null | grep '^{' > $OUT/r2_pcie_ceiling_${n}gpu.json
done
python bench.py --steps 10 --warmup 3 --fm 0 --locate 0 --c1 0 --decode 0 2>/dev/null | grep '^{' > $OUT/r2_bench_c2_1gpu_short.json
for n in 2 4 8; do
  [ $n -le $NG ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n \
      bench.py --gpus $n --steps 10 --warmup 3 --fm 0 --locate 0 --c1 0 --decode 0 2>/dev/null | grep '^{' > $OUT/r2_bench_c2_${n}gpu.json
done
for f in $OUT/*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
except Exception as e:
    print("unreadable", e); sys.exit(0)
if d.get("tool")=="pcie_ceiling":
    print(d["n_gpus"], "GPUs: ceiling", round(d["e2e_ceiling_text_MBps"]), "MB/s of text;", {k:(round(v["h2d_GBps"],1),round(v["d2h_GBps"],1)) for k,v in d["results"].items()})
else:
    print(d["n_gpus"], "GPUs: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "long", round(d["e2e"].get("long_call_MBps_this_rank",0)))
PY
done
