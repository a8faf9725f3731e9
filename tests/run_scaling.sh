#!/bin/bash
# Scaling table of the three workloads at N GPUs of one box: bash tests/run_scaling.sh N  (writes gpurun_out/r2_{train,export,projection}_${N}gpu.json)
N=$1
if [ "$N" = "1" ]; then T="python"; else T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$N"; fi
timeout 300 $T bench.py --gpus $N --steps 20 --warmup 5 --single-precision --no-device-batches --no-cpu-baseline > gpurun_out/r2_train_${N}gpu.json 2> gpurun_out/r2_train_${N}gpu.err
timeout 300 $T bench.py --gpus $N --workload export > gpurun_out/r2_export_${N}gpu.json 2> gpurun_out/r2_export_${N}gpu.err
timeout 300 $T bench.py --gpus $N --workload projection > gpurun_out/r2_projection_${N}gpu.json 2> gpurun_out/r2_projection_${N}gpu.err
python - <<PY
import json
for f in ("train","export","projection"):
    try:
        d=json.loads(open(f"gpurun_out/r2_{f}_${N}gpu.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["ms_per_step"],4), round(d["value"]), d.get("detail",{}).get("data_parallel"), d.get("detail",{}).get("replicas_identical"), round(d["e2e"]["value"]))
    except Exception as e: print(f, "ERR", e)
PY
