#!/bin/bash
# N GPUs of one box (N = number visible): one process on several devices (C-ABI), strong and weak scaling under torchrun
set -u
N=$(python -c "import torch; print(torch.cuda.device_count())")
O=gpurun_out/r02_n$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_inflate_parallel.py -m gpu -q -x --timeout 300 -k "several_devices" 2>&1 | tail -2
timeout 600 python bench.py --config multi --steps 3 > $O/multi.json 2> $O/multi.err; echo "multi rc=$?"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --scaling strong > $O/strong.json 2> $O/strong.err; echo "strong rc=$?"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > $O/weak.json 2> $O/weak.err; echo "weak rc=$?"
python - <<PY
import json
for f in ("weak","strong"):
    try:
        d=json.loads(open("$O/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], round(d["value"],2), round(d["e2e"]["value"],2), d["scaling"], d["ms_per_step"])
    except Exception as e: print(f,"ERR",e)
try:
    d=json.loads(open("$O/multi.json").read().strip().splitlines()[-1]); print("multi", [(r["devices"], round(r["gbs"],2), round(r["efficiency_vs_1"],2)) for r in d["rows"]])
except Exception as e: print("multi ERR", e)
PY
tail -2 $O/*.err
