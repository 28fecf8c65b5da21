"""GPU tuning sweep: per-kernel CUDA-event times of the bench workload for several dgr_set_tuning settings."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from dreamgaussian_b200 import _lib, multiview, scene
from dreamgaussian_b200.rasterizer import GaussianRasterizationSettings

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=100000); ap.add_argument("--res", type=int, default=800)
    ap.add_argument("--sh-degree", type=int, default=3); ap.add_argument("--opacity", default="trained")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--tunings", default="1,1,0;1,1,1;2,2,1;4,4,1;1,2,1;2,4,1")
    a = ap.parse_args()
    dev = torch.device("cuda", 0); lib = _lib.load()
    cloud = scene.make_cloud(a.points, a.sh_degree, seed=0, opacity=a.opacity, anisotropic=True)
    cams = scene.bench_views(8, a.res, a.res)
    t = lambda x: torch.tensor(np.asarray(x, np.float32), device=dev)
    params = {k: t(v) for k, v in cloud.items()}
    bg = t(np.ones(3, np.float32))
    settings = [GaussianRasterizationSettings(image_height=a.res, image_width=a.res, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg,
                scale_modifier=1.0, viewmatrix=t(c.world_view_transform), projmatrix=t(c.full_proj_transform), sh_degree=a.sh_degree,
                campos=t(c.camera_center), prefiltered=False, debug=False) for c in cams]
    rng = np.random.default_rng(17)
    up = (t(rng.normal(size=(3, a.res, a.res))), None, t(rng.normal(size=(1, a.res, a.res))))
    vsr = multiview.ViewShardedRasterizer(a.points, (a.sh_degree + 1) ** 2, dev)
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
    out = {}
    for tun in a.tunings.split(";"):
        pf, pb, od = (int(x) for x in tun.split(","))
        _lib.check(lib.dgr_set_tuning(pf, pb, od))
        for i in range(3):
            vsr.render_views(params, [settings[i % 8]], [up])
        torch.cuda.synchronize()
        evs = []
        for i in range(a.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); vsr.render_views(params, [settings[i % 8]], [up]); e1.record(); evs.append((e0, e1))
        torch.cuda.synchronize()
        step_ms = float(np.mean([x.elapsed_time(y) for x, y in evs]))
        lib.dgr_profile_enable(1)
        for i in range(a.steps):
            flush.zero_(); vsr.render_views(params, [settings[i % 8]], [up])
        torch.cuda.synchronize()
        kern = {}
        for name, ms in _lib.profile_collect():
            kern.setdefault(name, []).append(ms)
        lib.dgr_profile_enable(0)
        kern = {k: round(float(np.mean(v)) * 1e3, 1) for k, v in kern.items()}
        out[tun] = dict(step_us=round(step_ms * 1e3, 1), kernels_us=kern, sum_us=round(sum(kern.values()), 1))
        print(tun, json.dumps(out[tun]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune_%dk_%d_%s.json" % (a.points // 1000, a.res, a.opacity)), "w"), indent=1)

if __name__ == "__main__":
    main()
