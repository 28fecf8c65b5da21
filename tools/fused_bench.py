"""SURVEY §8 row f1 measurement: one stage-1 style render step (forward + backward + densification bookkeeping) of the
raw GaussianModel parameters, (a) in the reference's formulation — torch activations + torch.cat, the plain op, then the
three torch statistics updates (gs_renderer.py:196-216, 625-627; main.py:279-281) — against (b) FusedGaussianRasterizer,
which does all of it inside the per-Gaussian kernels.  Device-resident inputs, CUDA events per step, L2 flushed between
steps.  Prints one JSON line and writes gpurun_out/fused_bench.json."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from dreamgaussian_b200 import scene
from dreamgaussian_b200.fused import DensifyStats, FusedGaussianRasterizer
from dreamgaussian_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=100000); ap.add_argument("--res", type=int, default=800)
    ap.add_argument("--sh-degree", type=int, default=3); ap.add_argument("--steps", type=int, default=200)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    P = a.points
    cloud = scene.make_cloud(P, a.sh_degree, seed=0, opacity="trained", anisotropic=True)
    raw = {k: torch.tensor(v, device=dev).requires_grad_(True) for k, v in scene.to_raw_parameters(cloud).items()}
    cams = scene.bench_views(8, a.res, a.res)
    t = lambda x: torch.tensor(np.asarray(x, np.float32), device=dev)
    bg = t(np.ones(3, np.float32))
    settings = [GaussianRasterizationSettings(image_height=a.res, image_width=a.res, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg,
                scale_modifier=1.0, viewmatrix=t(c.world_view_transform), projmatrix=t(c.full_proj_transform), sh_degree=a.sh_degree,
                campos=t(c.camera_center), prefiltered=False, debug=False) for c in cams]
    rng = np.random.default_rng(17)
    gC, gA = t(rng.normal(size=(3, a.res, a.res))), t(rng.normal(size=(1, a.res, a.res)))
    stats = DensifyStats(P, dev)
    xyz_acc, denom, max_r = torch.zeros((P, 1), device=dev), torch.zeros((P, 1), device=dev), torch.zeros((P,), device=dev)
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)

    def zero():
        for v in raw.values():
            v.grad = None

    def reference_formulation(i):
        zero()
        m2d = torch.zeros_like(raw["xyz"], requires_grad=True)
        color, radii, depth, alpha = GaussianRasterizer(settings[i % 8])(
            means3D=raw["xyz"], means2D=m2d, shs=torch.cat((raw["features_dc"], raw["features_rest"]), dim=1),
            opacities=torch.sigmoid(raw["opacity"]), scales=torch.exp(raw["scaling"]),
            rotations=torch.nn.functional.normalize(raw["rotation"]))
        torch.autograd.backward([color, alpha], [gC, gA])
        with torch.no_grad():
            vis = radii > 0
            max_r[vis] = torch.max(max_r[vis], radii[vis].float())
            xyz_acc[vis] += torch.norm(m2d.grad[vis, :2], dim=-1, keepdim=True)
            denom[vis] += 1

    def fused(i):
        zero()
        color, radii, depth, alpha = FusedGaussianRasterizer(settings[i % 8])(
            raw["xyz"], raw["features_dc"], raw["features_rest"], raw["opacity"], raw["scaling"], raw["rotation"], stats=stats)
        torch.autograd.backward([color, alpha], [gC, gA])

    out = {"workload": "%dk Gaussians, %dx%d, SH deg %d, raw parameters, fwd+bwd+densify stats" % (P // 1000, a.res, a.res, a.sh_degree)}
    for name, fn in (("torch_activations_plus_plain_op", reference_formulation), ("fused", fused)):
        for i in range(10): fn(i)
        torch.cuda.synchronize()
        evs = []
        for i in range(a.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(i); e1.record(); evs.append((e0, e1))
        torch.cuda.synchronize()
        ms = np.array([x.elapsed_time(y) for x, y in evs])
        out[name] = {"ms_per_step_median": float(np.median(ms)), "ms_per_step_mean": float(ms.mean()),
                     "splats_per_s": float(P / (np.median(ms) * 1e-3))}
    out["speedup"] = out["torch_activations_plus_plain_op"]["ms_per_step_median"] / out["fused"]["ms_per_step_median"]
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fused_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
