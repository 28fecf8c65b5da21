"""SURVEY §8 row f4 measurement: extract_fields (resolution 128, 16^3 blocks — the reference defaults, gs_renderer.py:219)
of a 100k-Gaussian model: this library's kernels (CUDA events) against the CPU oracle (numpy restatement of the reference's
per-block loop) on the same inputs.  The reference's own torch implementation needs its un-installable imports; its
structure (a Python triple loop over 4096 blocks materialising [512 x L x 3] tensors) is what the oracle restates."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from dreamgaussian_b200 import _lib, fields, scene
from oracle import fields_oracle

out = {"rows": []}
for P, res in ((100000, 128), (500000, 128), (100000, 256)):
    raw = scene.to_raw_parameters(scene.make_cloud(P, 0, seed=4, sigma=0.0128 * (100000 / P) ** (1 / 3)))
    t = [torch.tensor(raw[k], device="cuda") for k in ("xyz", "opacity", "scaling", "rotation")]
    for _ in range(3): fields.extract_fields(*t, resolution=res)
    torch.cuda.synchronize()
    evs = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); occ, c, s = fields.extract_fields(*t, resolution=res); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    row = {"gaussians": P, "resolution": res, "ours_ms": float(np.median([a.elapsed_time(b) for a, b in evs]))}
    lib = _lib.load(); lib.dgr_profile_enable(1)
    fields.extract_fields(*t, resolution=res); torch.cuda.synchronize()
    row["kernels_ms"] = {k: round(v, 4) for k, v in _lib.profile_collect()}
    lib.dgr_profile_enable(0)
    if P == 100000 and res == 128:
        t0 = time.perf_counter()
        want, _, _ = fields_oracle.extract_fields(raw["xyz"], raw["opacity"], raw["scaling"], raw["rotation"], res, 16, 1.5)
        row["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
        row["max_abs_diff_vs_oracle"] = float(np.abs(occ.cpu().numpy() - want).max())
        row["speedup_vs_cpu_oracle"] = row["cpu_oracle_ms"] / row["ours_ms"]
    out["rows"].append(row); print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fields_bench.json"), "w"), indent=1)
