#!/bin/bash
# Round-2 final single-GPU measurement set (outputs under gpurun_out/, copied to profiles/ afterwards)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/r02_pytest_gpu.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
python bench.py > gpurun_out/r02_bench_1gpu_f64_final.json 2> gpurun_out/r02_bench_1gpu_f64_final.err
python bench.py --dtype f32 > gpurun_out/r02_bench_1gpu_f32_final.json 2> gpurun_out/r02_bench_1gpu_f32_final.err
python bench.py --workload se_atten --ncopy 14 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_se_atten_1gpu_f64_final.json 2> gpurun_out/r02_bench_se_atten_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_launches_final.log 2>&1
timeout 1200 ncu --set full --clock-control none --profile-from-start off -o /tmp/r02_step python tools/prof_step.py > gpurun_out/r02_prof_step.log 2>&1
python tools/ncu_step_summary.py /tmp/r02_step.ncu-rep 98304 gpurun_out/r02final > gpurun_out/r02_ncu_step_summary.log 2>&1
python tools/op_bench_refgpu.py > gpurun_out/r02_op_bench_refgpu_final.log 2>&1; tail -12 gpurun_out/r02_op_bench_refgpu_final.log
python - <<PY
import json
for f in ("r02_bench_1gpu_f64_final", "r02_bench_1gpu_f32_final", "r02_bench_se_atten_1gpu_f64_final", "r02_bench_reference_arm"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("unit"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"), (d.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print(f, "FAILED", e)
PY
