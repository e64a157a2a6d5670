#!/bin/bash
# strong-scaling runs of bench.py on the GPUs of this box (launched like the driver does)
O=gpurun_out; mkdir -p $O
NG=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  [ $n -gt $NG ] && break
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > $O/r1g_bench_n$n.json 2> $O/r1g_bench_n$n.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 50 --warmup 5 > $O/r1g_bench_n$n.json 2> $O/r1g_bench_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("$O/r1g_bench_n$n.json").read().strip().splitlines()[-1])
    print("N=$n", "ms", round(d["ms_per_step"],3), "b2b", round(d["config"]["ms_per_step_back_to_back_no_flush"],3), "e2e_ms", round(d["e2e"]["value"]*1e3,3), "verdict", d["config"]["verdict"])
except Exception as e:
    print("N=$n failed", e)
PY
done
