"""Per-kernel numbers of a `ncu --set full` capture as JSON for bench.py (profiles/r2_ncu_kernels.json): DRAM bytes and warp
instructions per launch, stamped with the hash of the kernel sources they were taken from (bench.py refuses a stale file).
Usage (here, no GPU needed): python tools/ncu_kernels_json.py gpurun_out/r2_prof.ncu-rep > profiles/r2_ncu_kernels.json"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dreamgaussian_b200 import build

NAMES = {"preprocess_fwd_kernel": "preprocess_fwd", "tile_scan_kernel": "tile_scan", "emit_instances_kernel": "emit_instances",
         "tile_sort_gather_kernel": "tile_sort_gather", "tile_sort_gather_big_kernel": "tile_sort_gather_big",
         "render_fwd_kernel": "render_fwd", "render_bwd_kernel": "render_bwd", "preprocess_bwd_kernel": "preprocess_bwd"}


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, m):
        v, u = float(r[idx[m]].replace(",", "")), units[idx[m]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    acc = {}
    for r in data:
        base = r[idx["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").replace("dgr::", "").strip()
        k = NAMES.get(base)
        if not k:
            continue
        a = acc.setdefault(k, dict(n=0, dram=0.0, inst=0.0, us=0.0, issue=0.0))
        a["n"] += 1
        a["dram"] += val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        a["inst"] += float(r[idx["smsp__inst_executed.sum"]].replace(",", ""))
        t = float(r[idx["gpu__time_duration.sum"]].replace(",", "")); a["us"] += t / 1000.0 if units[idx["gpu__time_duration.sum"]] in ("ns", "nsecond") else t
        a["issue"] += float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]].replace(",", ""))
    out = {"lib_source_hash": build.step_kernel_hash()[:16], "source": "ncu --set full --clock-control none, %s" % os.path.basename(rep),
           "workload": "100k Gaussians, 800x800, SH degree 3, forward+backward, opacity=trained, anisotropic",
           "dram_bytes_per_launch": {k: a["dram"] / a["n"] for k, a in acc.items()},
           "warp_inst_per_launch": {k: a["inst"] / a["n"] for k, a in acc.items()},
           "ncu_time_us": {k: a["us"] / a["n"] for k, a in acc.items()},
           "issue_active_pct": {k: a["issue"] / a["n"] for k, a in acc.items()},
           "launches_captured": {k: a["n"] for k, a in acc.items()}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
