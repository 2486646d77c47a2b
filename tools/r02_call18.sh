#!/bin/bash
# N GPUs of one box: strong scaling of one C3 + C2 batch (cut by b200z_partition_by_bytes), config C5 (reduced grid: --small) on N GPUs
set -u
N=$(python -c "import torch; print(torch.cuda.device_count())")
O=gpurun_out/r02_n$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --scaling strong > $O/strong.json 2> $O/strong.err; echo "strong rc=$?"
timeout 900 $TR --master-port 29513 bench.py --gpus $N --config c5 --small > $O/c5_small.json 2> $O/c5_small.err; echo "c5 rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$O/strong.json").read().strip().splitlines()[-1]); print("strong", d["n_gpus"], round(d["value"],2), round(d["e2e"]["value"],2), d["scaling"], d["ms_per_step"], d.get("handles"))
except Exception as e: print("strong ERR",e)
try:
    d=json.loads(open("$O/c5_small.json").read().strip().splitlines()[-1]); print("c5", d["n_gpus"], d["all_parity"], [(p["level"],p["size"],round(p["gbs"],1)) for p in d["points"]])
except Exception as e: print("c5 ERR", e)
PY
for f in $O/*.err; do tail -n 2 $f; done
