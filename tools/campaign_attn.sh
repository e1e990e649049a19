#!/bin/bash
# DPA-1 attention-layer measurement set on one GPU (outputs under gpurun_out/, copied to profiles/ afterwards).
# NOTE: the `ncu --set full` pass over a whole step replays ~600 launches ~40 times each: ~15 minutes of box time.
mkdir -p gpurun_out
python -m pytest tests/test_attn_layers.py -q -m gpu > gpurun_out/r02_pytest_attn.log 2>&1; tail -2 gpurun_out/r02_pytest_attn.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_attn_layers.py -q -m gpu -k "autograd or embed" > gpurun_out/r02_memcheck_attn.log 2>&1; echo "memcheck rc $?"; tail -4 gpurun_out/r02_memcheck_attn.log
python bench.py --workload dpa1_attn --ncopy 6 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_dpa1_attn.json 2> gpurun_out/r02_bench_dpa1_attn.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_dpa1_attn.csv python bench.py --workload dpa1_attn --ncopy 3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_launches_dpa1_attn.log 2>&1
NCOPY=3 MODEL=dpa1_attn timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/r02_attn_step python tools/prof_step.py > gpurun_out/r02_prof_attn_step.log 2>&1
python tools/ncu_step_summary.py /tmp/r02_attn_step.ncu-rep 5184 gpurun_out/r02attn > gpurun_out/r02_ncu_attn_step_summary.txt 2>&1; cat gpurun_out/r02_ncu_attn_step_summary.txt
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_dpa1_attn.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), d.get("roofline",{}).get("kernel"), d.get("roofline",{}).get("frac"), d["config"].get("attn_slots_evaluated_per_atom"))
for k,r in sorted(d.get("kernels",{}).items(), key=lambda kv:-kv[1]["ms_per_step"])[:12]:
    print("%-40s %8.3f ms share %s frac %s"%(k[:40], r["ms_per_step"], r.get("share"), r.get("frac")))
PY
