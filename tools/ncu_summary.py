"""Condense an ncu report (--page raw --csv) into the few metrics the roofline discussion needs."""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("sm__cycles_elapsed.avg", "cycles"),
    ("sm__cycles_elapsed.avg.per_second", "clk"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_tc%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("launch__registers_per_thread", "regs"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
]


def main(path, out=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines = ["kernel," + ",".join(f"{n}[{units[col[m]]}]" if m in col else n for m, n in WANT)]
    for r in rows[2:]:
        name = r[col["Kernel Name"]].replace(",", ";")[:70]
        vals = []
        for m, _ in WANT:
            v = r[col[m]] if m in col else ""
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            vals.append(v)
        lines.append(name + "," + ",".join(vals))
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
