#!/bin/bash
# bench.py under torchrun on N GPUs of one box (development aid; the driver runs the same command at round end)
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err
tail -3 gpurun_out/bench_r2_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2_${N}gpu.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "c4", json.dumps(d.get("c4"))[:1200])
PY
