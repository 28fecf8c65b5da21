"""Summarise an .ncu-rep (read here with `ncu -i`, no GPU needed) into a small markdown table for profiles/."""
import csv, io, subprocess, sys

KEYS = [("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_%"),
        ("smsp__inst_executed.sum", "warp_inst"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block")]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(k for _, k in KEYS) + " |")
    print("|---|" + "---|" * len(KEYS))
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("dgr::", "")[:44]
        vals = []
        for m, _ in KEYS:
            if m not in idx:
                vals.append("-"); continue
            v, u = r[idx[m]], units[idx[m]]
            try:
                f = float(v.replace(",", ""))
                if m.startswith("gpu__time"):
                    f = f / 1000.0 if u == "ns" else f
                    vals.append("%.1f" % f)
                elif "bytes" in m:
                    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                    vals.append("%.2f MB" % (f * scale))
                elif f == int(f) and abs(f) >= 1000:
                    vals.append("%d" % f)
                else:
                    vals.append("%.1f" % f)
            except ValueError:
                vals.append(v)
        print("| " + name + " | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
