"""Summarise an `ncu --set full` capture of ONE force step (tools/prof_step.py) per operator:
    python tools/ncu_step_summary.py gpurun_out/r02_step.ncu-rep <natoms> profiles/r02
writes <prefix>_ncu_full_summary.csv (one row per launch), <prefix>_dram_bytes_per_atom.json and
<prefix>_ncu_limiters.json (both read by bench.py)."""
import csv
import json
import subprocess
import sys

OPS = [("k_embed_grad", "se_atten_embed_grad"), ("k_embed", "se_atten_embed"), ("k_qkv_norm", "attn_qkv_normalize (+grad)"),
       ("k_attn_weights_grad", "attn_weights_grad"), ("k_attn_weights", "attn_weights"),
       ("k_res_ln_grad", "attn_residual_layernorm_grad"), ("k_res_ln", "attn_residual_layernorm"),
       ("k_rhat", "se_atten_rhat (+grad)"), ("gemm", "library GEMM (cuBLAS)"), ("k_gate_scalars", "se_atten_gate_scalars"),
       ("k_env_mat_a", "prod_env_mat_a"), ("k_tab_fwd", "tabulate_sections_desc"), ("k_tab_grad", "tabulate_sections_grad"),
       ("k_desc_bwd", "se_a_descriptor_grad"), ("k_fit_gemm", "fit_gemm_i8"), ("k_fit_slice<(int)6, (bool)1>", "fit_head"),
       ("k_fit_slice<(int)6, (bool)0>", "fit_slice_rows"), ("k_force_virial", "prod_force_virial_a"),
       ("k_split", "split glue"), ("k_mlp", "mlp glue")]
M = {
    "time_us": "gpu__time_duration.sum",
    "dram_rd": "dram__bytes_read.sum",
    "dram_wr": "dram__bytes_write.sum",
    "dram%": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex%": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts%": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor%": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "fp64%": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "issue%": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "warps%": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0,
         "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def main(rep, natoms, prefix):
    if rep.endswith(".csv"):  # already exported with `ncu -i x.ncu-rep --page raw --csv` (the report itself is large)
        with open(rep) as f:
            raw = f.read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, m):
        if m not in col:
            return None
        try:
            return float(r[col[m]]) * SCALE.get(units[col[m]], 1.0)
        except ValueError:
            return None

    out = [["operator", "kernel"] + list(M)]
    agg = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        op = next((o for k, o in OPS if k in name), "other")
        v = {k: val(r, m) for k, m in M.items()}
        out.append([op, name.replace(",", ";")[:90]] + [("" if v[k] is None else f"{v[k]:.6g}") for k in M])
        a = agg.setdefault(op, dict(time_us=0.0, bytes=0.0, lim=[]))
        a["time_us"] += v["time_us"] or 0.0
        a["bytes"] += (v["dram_rd"] or 0.0) + (v["dram_wr"] or 0.0)
        a["lim"].append((v["time_us"] or 0.0, v))
    with open(prefix + "_ncu_full_summary.csv", "w") as f:
        for r in out:
            f.write(",".join(map(str, r)) + "\n")
    dram = {op: a["bytes"] / natoms for op, a in agg.items()}
    lim = {}
    for op, a in agg.items():
        v = max(a["lim"], key=lambda x: x[0])[1]  # the longest launch of the operator speaks for it
        parts = [f"{k} {v[k]:.0f}" for k in ("dram%", "l1tex%", "lts%", "tensor%", "fp64%", "issue%", "warps%") if v.get(k) is not None and v[k] >= 5]
        lim[op] = ", ".join(parts) + f" (regs {v['regs']:.0f})"
    with open(prefix + "_dram_bytes_per_atom.json", "w") as f:
        json.dump({"source": f"ncu --set full of one step, {natoms} atoms, tools/prof_step.py", "f64": dram}, f, indent=1)
    with open(prefix + "_ncu_limiters.json", "w") as f:
        json.dump(lim, f, indent=1)
    tot = sum(a["time_us"] for a in agg.values())
    for op, a in sorted(agg.items(), key=lambda x: -x[1]["time_us"]):
        print(f"{op:28s} {a['time_us'] / 1e3:8.3f} ms  {100 * a['time_us'] / tot:5.1f}%  dram {dram[op]:9.0f} B/atom  | {lim[op]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3])
