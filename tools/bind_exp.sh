for mode in slice shared none; do
  echo "== TC_B200_BIND=$mode"
  TC_B200_BIND=$mode TC_BENCH_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 8 --steps 20 --warmup 3 --fm 0 --locate 0 --c1 0 --decode 0 2>&1 | grep -E '^\{|ms/step' | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],4))
    else: print(l.strip())
"
done
nproc; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; nvidia-smi topo -m | head -14
