"""SURVEY §8 row f3 measurement: distCUDA2 of this library against the reference's own simple_knn.cu (oracle/_ref, compiled
unmodified) and the CPU oracle, on the reference's init cloud (uniform ball, gs_renderer.py:694-702).  The reference
synchronises the device internally (two cudaMemcpy D2H, cudaMalloc/Free, thrust temporaries), so both are timed by wall
clock around a device synchronise; ours additionally by CUDA events.  Writes gpurun_out/knn_bench.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from simple_knn._C import distCUDA2
from oracle import knn_oracle

def ball(P, seed=0):
    rng = np.random.default_rng(seed)
    phi, ct, r = rng.random(P) * 2 * np.pi, rng.random(P) * 2 - 1, 0.5 * np.cbrt(rng.random(P))
    st = np.sqrt(1 - ct * ct)
    return np.ascontiguousarray(np.stack([r * st * np.cos(phi), r * st * np.sin(phi), r * ct], axis=1).astype(np.float32))

def wall(fn, reps):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3

out = {"unit": "ms per call (median)", "rows": []}
have_ref = os.path.exists(knn_oracle.REF_LIB)
for P in (5000, 100000, 1000000, 4000000):
    p = ball(P); t = torch.tensor(p, device="cuda")
    row = {"points": P, "ours_wall_ms": wall(lambda: distCUDA2(t), 20)}
    evs = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); distCUDA2(t); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    row["ours_event_ms"] = float(np.median([a.elapsed_time(b) for a, b in evs]))
    if have_ref:
        row["reference_wall_ms"] = wall(lambda: knn_oracle.reference_dist_cuda2(t), 10)
        row["speedup_vs_reference"] = row["reference_wall_ms"] / row["ours_wall_ms"]
        a, b = distCUDA2(t), knn_oracle.reference_dist_cuda2(t)
        row["max_rel_diff_vs_reference"] = float(((a - b).abs() / b.abs().clamp_min(1e-30)).max())
    if P <= 100000:
        t0 = time.perf_counter(); knn_oracle.dist2_f32(p); row["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
    row["points_per_s"] = P / (row["ours_event_ms"] * 1e-3)
    out["rows"].append(row); print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "knn_bench.json"), "w"), indent=1)
