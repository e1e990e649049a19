#!/usr/bin/env python
"""bench.py — us/step/atom of one compressed se_e2_a energy+force+virial evaluation.

    python bench.py --gpus N --steps K --warmup W            # this repository (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU ops on the host cores

N = 1: BASELINE.json configs[1] — the 192-atom water frame replicated 20x20x20 (1 536 000 atoms),
fp64, raw neighbour list with a 2 A skin rebuilt every 10 steps (reference MD set-up), formatted
every step.  N > 1: weak scaling, one 20x20x20 brick per GPU, spatial decomposition with a ghost
halo exchange over NCCL (one rank per GPU, launched by torchrun).

One JSON line on stdout (rank 0).  `value` = device-resident step, `e2e` = the same step through the
public host API (DeepPotB200.eval: pinned host coordinates in, host forces out).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "us/step/atom (energy+force+virial) compressed se_e2_a water"
UNIT = "us/step/atom"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--workload", default="water", choices=["water", "copper", "se_atten", "dpa1_attn"],
                    help="water: BASELINE config 2 (the metric's configuration); copper: config 3 (FCC, sel 512, rcut 8; "
                         "--ncopy = conventional cells per axis, 100 = 4 M atoms; 1 GPU only, evaluated in atom slabs); "
                         "se_atten: config 5 (DPA-1 se_atten_v2 strip / smooth, attn_layer 0, sel 120; --ncopy 14 = 526 848 atoms)")
    ap.add_argument("--ncopy", type=int, default=20, help="replicas of the 192-atom frame per axis and per GPU")
    ap.add_argument("--jitter", type=float, default=0.01)
    ap.add_argument("--cpu-ncopy", type=int, default=4, help="replicas per axis of the bounded CPU sample")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--atom-virial", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
                pw.append(float(p[3]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# DRAM traffic per atom and operator (dram__bytes_read.sum + dram__bytes_write.sum over the launches of one step):
# written by tools/ncu_step_summary.py from the `ncu --set full` capture of `bench.py --ncopy 8` committed under
# profiles/ (the kernels stream per-atom data, so the figure scales with the atom count).  No capture, no number.
def ncu_dram_bytes_per_atom():
    p = os.path.join(ROOT, "profiles", "r02_dram_bytes_per_atom.json")
    if not os.path.exists(p):
        return {}, None
    with open(p) as f:
        d = json.load(f)
    return d.get("f64", {}), "profiles/r02_dram_bytes_per_atom.json (" + d.get("source", "ncu --set full") + ")"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def cpu_pipeline_timing(args, steps, warmup, seconds=None):
    """The reference's CPU ops (OpenMP, all host cores) on a bounded sample of the workload."""
    import torch

    import __graft_entry__ as g
    from oracle import cpu as ocpu
    from oracle import pipeline
    from oracle.refmodel import RefWaterModel  # same seeds / tables as the product's SeAModel, no product import

    kind = "reference" if ocpu.available("reference") else "port"
    lib = ocpu.CpuLib(kind)
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to its workers)
    ncores = os.cpu_count() or 1
    try:
        import ctypes

        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(ncores)
    except OSError:
        pass
    torch.set_num_threads(ncores)
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    np_dt = np.float64 if args.dtype == "f64" else np.float32
    model = RefWaterModel(dtype)
    cfg = model.cfg
    coord, atype, box = g.water_box(args.cpu_ncopy, args.jitter)
    nat = len(atype)
    t0 = time.perf_counter()
    lists = pipeline.build_lists(ocpu.CpuLib("port"), coord.astype(np_dt), atype, box.astype(np_dt), cfg.rcut + 2.0)
    t_build = time.perf_counter() - t0
    for _ in range(warmup):
        pipeline.evaluate(lib, model, lists)
    stage = {}
    t0 = time.perf_counter()
    n = 0
    while n < steps:
        pipeline.evaluate(lib, model, lists, stage)
        n += 1
        if seconds is not None and time.perf_counter() - t0 > seconds and n >= 3:
            break
    dt = (time.perf_counter() - t0) / n
    us = dt * 1e6 / nat
    return dict(us_per_step_atom=us, ms_per_step=dt * 1e3, steps=n, natoms=nat, kind=kind,
                cores=os.cpu_count(), build_s=t_build,
                stages_us_per_atom={k: v / n * 1e6 / nat for k, v in stage.items()},
                sample=f"{nat}-atom water box ({args.cpu_ncopy}^3 replicas of the 192-atom frame), {n} steps, "
                       f"raw list (cell list, {t_build * 1e3:.0f} ms) excluded, {args.dtype}")


def cpu_attn_timing(args, steps, warmup, seconds=None):
    """DPA-1 with attention layers on the host cores: the reference's CPU env-mat / force / virial ops + a plain torch
    restatement of its dense uncompressed path (embedding MLP, attention layers, autograd) = oracle.pipeline_atten
    .evaluate_layers, on a bounded sample (the reference cannot compress this model, so this IS its algorithm)."""
    import torch

    import __graft_entry__ as g
    from oracle import cpu as ocpu
    from oracle import pipeline, pipeline_atten
    from oracle.refmodel import RefAttnModel

    kind = "reference" if ocpu.available("reference") else "port"
    lib = ocpu.CpuLib(kind)
    ncores = os.cpu_count() or 1
    try:
        import ctypes

        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(ncores)
    except OSError:
        pass
    torch.set_num_threads(ncores)
    model = RefAttnModel(torch.float64)
    ncopy = min(args.cpu_ncopy, 2)
    coord, atype, box = g.water_box(ncopy, args.jitter)
    nat = len(atype)
    lists = pipeline.build_lists(ocpu.CpuLib("port"), coord, atype, box, model.cfg.rcut + 2.0)
    for _ in range(warmup):
        pipeline_atten.evaluate_layers(lib, model, lists)
    t0 = time.perf_counter()
    n = 0
    while n < steps:
        pipeline_atten.evaluate_layers(lib, model, lists)
        n += 1
        if seconds is not None and time.perf_counter() - t0 > seconds and n >= 2:
            break
    dt = (time.perf_counter() - t0) / n
    return dict(us_per_step_atom=dt * 1e6 / nat, ms_per_step=dt * 1e3, steps=n, natoms=nat,
                kind=kind + " CPU ops + torch-CPU restatement of the dense attention path (autograd)", cores=ncores,
                stages_us_per_atom={},
                sample=f"{nat}-atom water box ({ncopy}^3 replicas of the 192-atom frame), {n} steps, raw list excluded, f64")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "dpa1_attn":
        r = cpu_attn_timing(args, args.steps, args.warmup)
        nat = 192 * args.ncopy ** 3 * args.gpus
        line = {
            "impl": "reference", "metric": METRIC, "value": r["us_per_step_atom"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": False,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": (f"DPA-1 se_atten_v2 with attention (strip, smooth, attn_layer 2, attn 128, dotr, sel "
                                    f"120; tabulated embedding + attention layers) water, {nat}-atom box ({args.ncopy}^3 "
                                    f"replicas of the 192-atom frame per GPU, jitter {args.jitter} A), f64, "
                                    f"{args.gpus}xB200"),
                       "natoms": nat, "rcut": 6.0, "rcut_smth": 0.5, "sel": [120], "neuron": [25, 50, 100],
                       "axis_neuron": 16, "fitting_neuron": [240, 240, 240], "bench_workload": "dpa1_attn",
                       "reference_implementation": "uncompressed dense path (the reference cannot compress a model "
                                                   "with attention layers) on all host cores; bounded sample",
                       "sample_natoms": r["natoms"]},
            "cpu_baseline": {"value": r["us_per_step_atom"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                             "sample": r["sample"]},
            "e2e": {"value": r["us_per_step_atom"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "product_native_loaded": _maps_contain("libdpb200"),
        }
        _emit(json.dumps(line))
        return
    r = cpu_pipeline_timing(args, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["us_per_step_atom"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": False,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        # the arm's own configuration (same workload string and model keys as run_ours), plus what was actually timed
        "config": {"workload": f"se_e2_a compressed water, {192 * args.ncopy ** 3 * args.gpus}-atom box ({args.ncopy}^3 "
                               f"replicas of the 192-atom frame per GPU, Gaussian jitter {args.jitter} A), {args.dtype}, "
                               f"{args.gpus}xB200",
                   "natoms": 192 * args.ncopy ** 3 * args.gpus, "rcut": 6.0, "rcut_smth": 0.5, "sel": [46, 92],
                   "neuron": [25, 50, 100], "axis_neuron": 16, "fitting_neuron": [240, 240, 240],
                   "table": "dp-compress restatement, stride 0.01/0.1, extrapolate 5, random-init weights (seed 1)",
                   "reference_implementation": "reference CPU ops (libdeepmd *_cpu, OpenMP, all host cores) + torch CPU "
                                               "fitting net; each step = one evaluation of a bounded sample of the "
                                               "workload (us/step/atom is size-independent)",
                   "sample_natoms": r["natoms"]},
        "cpu_baseline": {"value": r["us_per_step_atom"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"], "stages_us_per_atom": r["stages_us_per_atom"]},
        "e2e": {"value": r["us_per_step_atom"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "product_native_loaded": _maps_contain("libdpb200"),
    }
    _emit(json.dumps(line))


def _maps_contain(name):
    try:
        with open("/proc/self/maps") as f:
            return any(name in ln for ln in f)
    except OSError:
        return None


def _emit(text):
    import builtins

    getattr(builtins, "_dpb200_emit", print)(text)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime

        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    pkg = g.load_package()
    from deepmd_kit_b200 import ops
    from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

    L = pkg._lib.lib()
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    np_dt = np.float64 if args.dtype == "f64" else np.float32
    if args.workload == "copper":
        from deepmd_kit_b200.model import COPPER_CONFIG

        cfg = SeAConfig(**COPPER_CONFIG)
    elif args.workload == "se_atten":
        from deepmd_kit_b200.atten import SeAttenConfig, SeAttenModel

        cfg = SeAttenConfig()
    elif args.workload == "dpa1_attn":  # DPA-1 with its attention layers: not compressible upstream
        from deepmd_kit_b200.atten import SeAttenConfig, SeAttenModel

        cfg = SeAttenConfig(attn_layer=2)
    else:
        cfg = SeAConfig()
    model = SeAttenModel(cfg, dtype, dev) if args.workload in ("se_atten", "dpa1_attn") else SeAModel(cfg, dtype, dev)
    esz = 8 if args.dtype == "f64" else 4

    if world == 1:
        if args.workload == "copper":
            coord, atype, box = g.copper_box(args.ncopy, 0.05)
        else:
            coord, atype, box = g.water_box(args.ncopy, args.jitter)
        natoms_total = len(atype)
        dp = DeepPotB200(model, skin=2.0, nlist_every=10)
        coord_d = torch.as_tensor(coord.astype(np_dt)).to(dev)
        atype_d = torch.as_tensor(atype).to(dev)

        def step():
            return dp.eval_device(coord_d, atype_d, box, atom_virial=args.atom_virial)

        parallelism = "1 GPU"
    else:
        from deepmd_kit_b200.domain import DomainDeepPot, proc_grid

        grid = proc_grid(world)
        dp = DomainDeepPot(model, grid, skin=2.0, nlist_every=10)
        if args.workload == "copper":
            # strong scaling: ONE global box of ncopy^3 FCC cells, split into bricks
            coord, atype, box = dp.make_local_copper(args.ncopy, 0.05)
            nat_t = torch.tensor([len(atype)], dtype=torch.int64, device=dev)
            dist.all_reduce(nat_t)
            natoms_total = int(nat_t.item())
        else:
            coord, atype, box = dp.make_local_water(g.water_box, args.ncopy, args.jitter)
            natoms_total = len(atype) * world
        coord_d = torch.as_tensor(coord.astype(np_dt)).to(dev)
        atype_d = torch.as_tensor(atype).to(dev)

        def step():
            return dp.eval_device(coord_d, atype_d, box, atom_virial=args.atom_virial)

        parallelism = f"spatial {grid[0]}x{grid[1]}x{grid[2]} bricks, NCCL halo"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = L.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    marks = []
    for _ in range(args.steps):
        out = step()
        m = torch.cuda.Event(enable_timing=True)
        m.record()
        marks.append(m)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    step_ms = [round((ev0 if k == 0 else marks[k - 1]).elapsed_time(marks[k]), 3) for k in range(len(marks))]
    launches = L.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = ms_per_step * 1e3 / natoms_total
    energy = float(out[0])

    # ---- e2e through the public host API ---------------------------------------------------
    e2e = None
    if not args.no_e2e:
        nat_local = len(atype)
        k = max(3, min(args.steps, 10))
        coords_h = coord.astype(np_dt).reshape(1, -1)
        cells_h = np.asarray(box, np.float64).reshape(1, 9)
        for _ in range(2):
            dp.eval(coords_h, cells_h, atype)
        barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            res = dp.eval(coords_h, cells_h, atype)
        barrier()
        dt = (time.perf_counter() - t0) / k
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": dt * 1e6 / natoms_total, "unit": UNIT, "h2d_bytes_per_step": natoms_total * 3 * esz,
               "d2h_bytes_per_step": (natoms_total * 3 + 10 * world) * esz, "steps": k,
               "api": "DeepPotB200.eval(coords, cells, atom_types) — pinned host buffers, H2D + D2H inside the timed region"}

    # ---- per-kernel timing (instrumented pass, rank 0 reports) ---------------------------------
    # (every rank runs the instrumented steps -- they contain collectives -- rank 0 reports)
    kernels, roofline = per_kernel(args, torch, ops, model, dp, step, L, dev, len(atype), esz)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload in ("water", "dpa1_attn"):
        try:
            r = (cpu_pipeline_timing if args.workload == "water" else cpu_attn_timing)(args, 10 ** 6, 1,
                                                                                     seconds=args.cpu_seconds)
            cpu_base = {"value": r["us_per_step_atom"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                        "sample": r["sample"], "stages_us_per_atom": r["stages_us_per_atom"]}
        except Exception as e:  # the checker libraries are test infrastructure; report, do not hide
            cpu_base = {"value": None, "unit": UNIT, "error": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": False,
            "scaling": "strong" if args.workload == "copper" else "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {
                "workload": (f"se_e2_a compressed water, {natoms_total}-atom box ({args.ncopy}^3 replicas of the 192-atom "
                             f"frame per GPU, Gaussian jitter {args.jitter} A), {args.dtype}, {world}xB200")
                if args.workload == "water" else
                (f"DPA-1 se_atten_v2 (strip, smooth, attn_layer 0, sel 120) compressed water, {natoms_total}-atom box "
                 f"({args.ncopy}^3 replicas of the 192-atom frame per GPU, jitter {args.jitter} A), {args.dtype}, {world}xB200")
                if args.workload == "se_atten" else
                (f"DPA-1 se_atten_v2 with attention (strip, smooth, attn_layer 2, attn 128, dotr, sel 120; tabulated "
                 f"embedding + attention layers) water, {natoms_total}-atom box ({args.ncopy}^3 replicas of the 192-atom "
                 f"frame per GPU, jitter {args.jitter} A), {args.dtype}, {world}xB200")
                if args.workload == "dpa1_attn" else
                (f"se_e2_a compressed copper FCC, {natoms_total} atoms ({args.ncopy}^3 cells, a0 3.615 A, jitter 0.05 A), "
                 f"{args.dtype}, {world}xB200, evaluated in "
                 f"{1 if dp.state.chunks is None else len(dp.state.chunks)} atom slab(s) per GPU"),
                "natoms": natoms_total, "rcut": cfg.rcut, "rcut_smth": cfg.rcut_smth, "sel": list(cfg.sel),
                "neuron": list(cfg.neuron), "axis_neuron": cfg.axis_neuron, "fitting_neuron": list(cfg.fitting_neuron),
                "bench_workload": args.workload,
                "table": "dp-compress restatement, stride 0.01/0.1, extrapolate 5, random-init weights (seed 1)",
                "skin": 2.0, "nlist_every": 10, "parallelism": parallelism,
                "cuda_graph": bool(getattr(dp, "use_graph", False) and getattr(model, "graph_safe", True)),
                "atom_virial": bool(args.atom_virial),
                "l2_policy": "per-step working set (tens of GB of env-mat intermediates) is far larger than the 126 MB L2",
                "energy": energy,
                **({"attn_slots_evaluated_per_atom": float(np.mean(model.last_n_eff))}
                   if getattr(model, "last_n_eff", None) else {}),
            },
            "clocks": clocks, "gpu_launches": int(launches), "step_ms": step_ms,
        }
        if e2e:
            line["e2e"] = e2e
        if roofline:
            line["roofline"] = roofline
        if kernels:
            line["kernels"] = kernels
        if cpu_base:
            line["cpu_baseline"] = cpu_base
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def per_kernel(args, torch, ops, model, dp, step, L, dev, nloc, esz):
    """CUDA-event time of every dpb200 operator inside the real step (events on the launching stream),
    and the roofline of each against measured peaks."""
    names = ["prod_env_mat_a", "tabulate_sections_fwd", "tabulate_sections_desc", "tabulate_sections_grad",
             "prod_force_virial_a", "prod_force_virial_a_ex", "use_nlist_map",
             "normalize_coord", "copy_coord", "build_nlist", "se_a_descriptor", "se_a_descriptor_grad", "halo_pack",
             "halo_unpack_add", "fit_gemm_i8", "fit_head", "fit_slice_rows", "tabulate_fusion_se_a",
             "tabulate_fusion_se_a_grad", "split_i8_rows", "tabulate_fusion_se_atten_gate",
             "tabulate_fusion_se_atten_gate_grad", "tabulate_fusion_se_atten_gate_desc", "fit_slice_cols",
             "se_atten_gate_scalars", "prod_force_virial_a_pair", "se_atten_embed", "se_atten_embed_grad",
             "se_atten_rhat", "se_atten_rhat_grad", "attn_qkv_normalize", "attn_qkv_normalize_grad", "attn_weights",
             "attn_weights_grad", "attn_residual_layernorm", "attn_residual_layernorm_grad"]
    acc = {}
    orig = {}

    def wrap(name, fn):
        def f(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            acc.setdefault(name, []).append((e0, e1))
            return r
        return f

    for n in names:
        orig[n] = getattr(ops, n)
        setattr(ops, n, wrap(n, orig[n]))
    umbrella = "fitting_net (total: fit_gemm_i8 + fit_head + fit_slice_rows + se_a_descriptor_grad + torch glue)"
    has_fit = hasattr(model, "energy_and_dy")
    if has_fit:
        orig_fit = model.energy_and_dy
        orig_fit_split = model.energy_and_dy_split
        model.energy_and_dy = wrap(umbrella, orig_fit)
        model.energy_and_dy_split = wrap(umbrella, orig_fit_split)
    # attention model: the dense products of the layers run on the library; one row for all of them
    from deepmd_kit_b200 import atten as atten_mod

    gemm_row = "attn_library_gemm (cuBLAS fp64 / fp32 through torch: in_proj, q k^T, A v, out_proj and transposes)"
    gemm_orig = {n: getattr(atten_mod, n) for n in ("_addmm", "_bmm", "_mm")}
    if args.workload == "dpa1_attn":
        for n, fn in gemm_orig.items():
            setattr(atten_mod, n, wrap(gemm_row, fn))
    nsteps = 10 if args.workload != "dpa1_attn" else 3
    graph_mode = getattr(dp, "use_graph", False)
    dp.use_graph = False  # the instrumented pass needs real launches (events cannot be timed inside a graph)
    try:
        for _ in range(2):  # eager warm-up: the graph capture emptied the allocator cache
            out = step()
        torch.cuda.synchronize()
        acc.clear()
        for _ in range(nsteps):
            out = step()
        torch.cuda.synchronize()
    finally:
        for n in names:
            setattr(ops, n, orig[n])
        for n, fn in gemm_orig.items():
            setattr(atten_mod, n, fn)
        if has_fit:
            model.energy_and_dy = orig_fit
            model.energy_and_dy_split = orig_fit_split
        dp.use_graph = graph_mode
    nlist = out[3]["nlist"]
    cfg = model.cfg
    if nlist is None:  # slab-wise evaluation: count the real neighbours of the first slab
        st0 = dp.state
        a, b = st0.chunks[0][0], st0.chunks[0][1]
        nl0 = orig["prod_env_mat_a"](dp._last_ext_coord.reshape(-1), st0.ext_type, st0.numneigh, st0.rows, model.davg,
                                     model.dstd, st0.nloc, int(st0.ext_type.numel()), cfg.rcut, cfg.rcut_smth, cfg.sec,
                                     row_range=(a, b))[3]
        nreal = float((nl0 >= 0).sum().item()) / (b - a)
        del nl0
    else:
        nreal = float((nlist >= 0).sum().item()) / nloc
    nnei, M, nt = cfg.nnei, model.M, len(cfg.sel)  # nt = tables per atom (type sections)
    st = dp.state
    nall = int(st.ext_type.numel())
    raw = float(st.numneigh.sum().item()) / nloc
    F = esz
    peaks, hbm_src = measured_peaks()
    hbm = float(peaks["hbm_gbs"])
    fma = L.fma_peak(args.dtype, None)
    npr = nreal + nt  # one folded padding entry per table
    # int8 tensor-core work of the fitting net: every GEMM of the forward and of the input-gradient backward as
    # NS(NS+1)/2 exact slice products (csrc/fit_tc.cu; NS = 6 for the fp64 model: 21, NS = 4 for fp32: 10), 2 ops per MAC
    widths = [getattr(model, "dim_in", M * cfg.axis_neuron)] + list(cfg.fitting_neuron)
    fit_mac = sum(a * b for a, b in zip(widths[:-1], widths[1:]))
    fit_ns = 6 if args.dtype == "f64" else 4
    fit_int8_ops = 2.0 * (fit_ns * (fit_ns + 1) // 2) * (2 * fit_mac)
    int8_peak = 2.0 * float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    alg = {
        "prod_env_mat_a": ("hbm", (19 * nnei * F + 4 * nnei) + 4 * raw + 3 * F * (1 + nall / nloc)),
        "prod_force_virial_a": ("hbm", (19 * nnei * F + 4 * nnei) + 3 * F + (9 * F if args.atom_virial else 0)),
        "prod_force_virial_a_ex": ("hbm", (19 * nnei * F + 4 * nnei) + 3 * F + (9 * F if args.atom_virial else 0)),
        "tabulate_sections_fwd": ("fp", 18 * npr * M),
        "tabulate_sections_desc": ("fp", 18 * npr * M),  # the fused descriptor epilogue is not counted
        "tabulate_sections_grad": ("fp", 36 * npr * M),
        # descriptor backward: reads dE/dD [M*axis] and GR [4*M], writes dE/dGR [4*M]
        "se_a_descriptor_grad": ("hbm", (M * cfg.axis_neuron + 8 * M) * F),
        "fit_gemm_i8": ("tensor", fit_int8_ops),
    }
    if args.workload == "se_atten":
        # the gated operator streams the materialised two_embed rows of the real neighbours (reference op signature)
        alg["tabulate_fusion_se_a"] = ("hbm", nreal * M * F + 5 * nnei * F + 4 * M * F)
        alg["tabulate_fusion_se_a_grad"] = ("hbm", 2 * nreal * M * F + 10 * nnei * F + 4 * M * F)
        # pair-indexed gate: no two_embed stream, the op is bound like the plain se_a table kernels
        alg["tabulate_fusion_se_atten_gate"] = ("fp", 20 * npr * M)
        alg["tabulate_fusion_se_atten_gate_desc"] = ("fp", 20 * npr * M)  # the fused descriptor epilogue is not counted
        alg["prod_force_virial_a_pair"] = ("hbm", (19 * nnei * F + 4 * nnei) + 2 * nnei * F + 3 * F)
        alg["se_atten_gate_scalars"] = ("hbm", nnei * (4 + 3 * F) + nnei * (4 + 2 * F))
        alg["tabulate_fusion_se_atten_gate_grad"] = ("fp", 42 * npr * M)
    if args.workload == "dpa1_attn":
        nl_, h_ = cfg.attn_layer, cfg.attn
        # slots actually evaluated per atom: trailing empty slots of a slab are folded into one (atten.py attn_compact)
        ne_ = float(np.mean(getattr(model, "last_n_eff", [nnei]) or [nnei]))
        alg["se_atten_embed"] = ("hbm", 3 * ne_ * M * F + ne_ * (2 * F + 4))
        alg["se_atten_embed_grad"] = ("hbm", 3 * ne_ * M * F + ne_ * (4 * F + 4))
        alg["attn_qkv_normalize"] = ("hbm", nl_ * 2 * ne_ * 3 * h_ * F)
        alg["attn_qkv_normalize_grad"] = ("hbm", nl_ * 3 * ne_ * 3 * h_ * F)
        alg["attn_weights"] = ("hbm", nl_ * 3 * ne_ * ne_ * F)
        alg["attn_weights_grad"] = ("hbm", nl_ * 4 * ne_ * ne_ * F)
        alg["attn_residual_layernorm"] = ("hbm", nl_ * 4 * ne_ * M * F)
        alg["attn_residual_layernorm_grad"] = ("hbm", nl_ * 3 * ne_ * M * F)
        alg["prod_force_virial_a_pair"] = ("hbm", (19 * nnei * F + 4 * nnei) + 2 * nnei * F + 3 * F)
        alg["se_atten_gate_scalars"] = ("hbm", nnei * (4 + 3 * F) + nnei * (4 + 2 * F))
        # forward products of a layer; the backward has two products per forward one
        alg[gemm_row] = ("fp", nl_ * 3 * 2.0 * ne_ * (M * 3 * h_ + 2 * ne_ * h_ + h_ * M))
    table = {}
    total = 0.0
    for n, evs in acc.items():
        times = sorted(a.elapsed_time(b) for a, b in evs)
        med = times[len(times) // 2]
        # median call x calls per step (robust against allocator hiccups); kernels with several call shapes per
        # step (the six GEMMs of the fitting net) are summed instead
        ms = sum(times) / nsteps if n.startswith(("fit_", "attn_library")) else med * len(evs) / nsteps
        table[n] = {"ms_per_step": ms, "calls_per_step": len(evs) / nsteps, "ms_min_call": times[0],
                    "ms_max_call": times[-1]}
        if n == umbrella:
            table[n]["umbrella"] = True  # contains other rows of this table: not part of the shares
        else:
            total += ms
    for n, row in table.items():
        row["share"] = None if row.get("umbrella") else (row["ms_per_step"] / total if total > 0 else None)
        if n in alg and row["ms_per_step"] > 0:
            kind, per_atom = alg[n]
            t = row["ms_per_step"] * 1e-3
            if kind == "tensor":
                a = per_atom * nloc / t / 1e12
                row.update(bound="tensor", achieved=a, peak=int8_peak, unit="TOP/s (int8)", frac=a / int8_peak,
                           alg_int8_ops_per_atom=per_atom, fp64_equivalent_tflops=2.0 * 2 * fit_mac * nloc / t / 1e12,
                           peak_source=hbm_src + ": 2 x bf16_tflops_sustained (dense int8 runs at twice the bf16 rate on "
                                                 "B200; the kernel is timed inside a long step)")
            elif kind == "hbm":
                a = per_atom * nloc / t / 1e9
                row.update(bound="hbm", achieved=a, peak=hbm, unit="GB/s", frac=a / hbm, alg_bytes_per_atom=per_atom)
            else:
                a = per_atom * nloc / t / 1e12
                row.update(bound="fp64-fma" if args.dtype == "f64" else "fp32-fma", achieved=a, peak=fma,
                           unit="TFLOP/s", frac=a / fma, alg_flops_per_atom=per_atom)
    # what ncu names as the limiter of each kernel: profiles/r02_ncu_full_summary.csv (tools/ncu_step_summary.py)
    lim_path = os.path.join(ROOT, "profiles", "r02_ncu_limiters.json")
    if args.dtype == "f64" and os.path.exists(lim_path):
        with open(lim_path) as f:
            limiter = json.load(f)
        for n, row in table.items():
            if n in limiter:
                row["ncu_limiter"] = limiter[n]
    ours = {n: r for n, r in table.items() if "bound" in r and not n.startswith("attn_library")}  # (library row: reported, not ours)
    top = max(ours, key=lambda n: ours[n]["ms_per_step"]) if ours else None
    roofline = None
    if top:
        r = ours[top]
        dram, dram_src = ncu_dram_bytes_per_atom()
        traffic = dram[top] * nloc if (args.dtype == "f64" and top in dram) else None
        roofline = {"kernel": top, "bound": "tensor" if r["bound"] == "tensor" else ("hbm" if r["bound"] == "hbm" else r["bound"]),
                    "achieved": r["achieved"], "peak": r["peak"], "unit": r["unit"], "frac": r["frac"], "traffic": traffic,
                    "traffic_note": ("DRAM bytes per step of this operator, " + dram_src) if traffic is not None else
                                    "no ncu capture of this kernel is committed for this build",
                    "peak_source": r.get("peak_source", hbm_src if r["bound"] == "hbm" else
                                         "dpb200_fma_peak measured in this run (burst); nominal B200 FP64 is ~40 TFLOP/s"),
                    "mean_real_neighbours": nreal, "mean_raw_neighbours": raw}
        # every operator with a roofline, largest first (the headline object above is the dominant one)
        roofline["all"] = [dict(kernel=n, ms_per_step=ours[n]["ms_per_step"], bound=ours[n]["bound"],
                                achieved=ours[n]["achieved"], peak=ours[n]["peak"], unit=ours[n]["unit"],
                                frac=ours[n]["frac"])
                           for n in sorted(ours, key=lambda q: -ours[q]["ms_per_step"])]
    return table, roofline


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: everything printed before it (NCCL's version banner, library
    # chatter) is sent to stderr by pointing fd 1 at fd 2 until the result is ready
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved, "w")
    import builtins

    def emit(text):
        real_stdout.write(text + "\n")
        real_stdout.flush()

    builtins._dpb200_emit = emit
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
