#!/bin/bash
# usage: bash tools/campaign_multi.sh N  -- round-2 multi-GPU measurement set on N GPUs of one box (outputs in gpurun_out/)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { if [ "$N" = "1" ]; then timeout 600 python bench.py "$@"; else timeout 600 $TR --master-port $PORT bench.py "$@"; fi; }
PORT=29611 run --gpus $N --workload copper --ncopy 100 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_copper4m_${N}gpu.json 2> gpurun_out/r02_copper4m_${N}gpu.err
tail -c 400 gpurun_out/r02_copper4m_${N}gpu.err
PORT=29612 run --gpus $N --workload se_atten --ncopy 14 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_se_atten_${N}gpu.json 2> gpurun_out/r02_se_atten_${N}gpu.err
tail -c 400 gpurun_out/r02_se_atten_${N}gpu.err
if [ "$N" != "1" ]; then timeout 500 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "$N" 2>&1 | tail -3; fi
python - <<PY
import json
for w in ("copper4m", "se_atten"):
    try:
        d = json.loads(open(f"gpurun_out/r02_{w}_${N}gpu.json").read().strip().splitlines()[-1])
        print(w, d["n_gpus"], d["value"], d["unit"], d["ms_per_step"], d.get("scaling"), (d.get("e2e") or {}).get("value"), d["config"].get("workload"))
    except Exception as e:
        print(w, "FAILED", e)
PY
