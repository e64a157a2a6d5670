#!/bin/bash
# round 2: multi-GPU bench line (launched like the driver does), N = $1
N=${1:-2}
O=gpurun_out; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --steps 20 --warmup 5 > $O/r2z_bench_n$N.json 2> $O/r2z_bench_n$N.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/r2z_bench_n$N.json").read().strip().splitlines() if l.startswith("{")][-1])
    c4 = d["config"].get("c4_subcube") or {}
    print("N=$N ms %.4f e2e %.4f slices/rank %s launches %d verdict %s graph %s c4_ms %.2f c4 %s" % (d["ms_per_step"], d["e2e"]["value"] * 1e3, d["config"]["slices_per_rank"], d["gpu_launches"], d["config"]["verdict"], d["config"]["cuda_graph"], c4.get("value", 0) * 1e3, c4.get("result")))
except Exception as e:
    print("FAILED", e); print(open("$O/r2z_bench_n$N.err").read()[-1500:])
PY
