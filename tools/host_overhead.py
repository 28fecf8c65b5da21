"""Host-side (Python / ctypes / launch) cost per fwd+bwd step of the two public entry points, with a cProfile breakdown.
The GPU work is ~0.29 ms per step at the bench shape; anything above that in wall time is host overhead."""
import argparse, cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from dreamgaussian_b200 import multiview, scene
from dreamgaussian_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=100000); ap.add_argument("--res", type=int, default=800)
    ap.add_argument("--steps", type=int, default=300); ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    cloud = scene.make_cloud(a.points, 3, seed=0, opacity="trained", anisotropic=True)
    cams = scene.bench_views(8, a.res, a.res)
    t = lambda x: torch.tensor(np.asarray(x, np.float32), device=dev)
    params = {k: t(v) for k, v in cloud.items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    bg = t(np.ones(3, np.float32))
    settings = [GaussianRasterizationSettings(image_height=a.res, image_width=a.res, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg,
                scale_modifier=1.0, viewmatrix=t(c.world_view_transform), projmatrix=t(c.full_proj_transform), sh_degree=3,
                campos=t(c.camera_center), prefiltered=False, debug=False) for c in cams]
    rng = np.random.default_rng(17)
    gC, gA = t(rng.normal(size=(3, a.res, a.res))), t(rng.normal(size=(1, a.res, a.res)))
    m2d = torch.zeros((a.points, 3), device=dev)
    vsr = multiview.ViewShardedRasterizer(a.points, 16, dev)

    def autograd_step(i):
        for v in leaves.values():
            v.grad = None
        color, radii, depth, alpha = GaussianRasterizer(raster_settings=settings[i % 8])(means2D=m2d, **leaves)
        loss = (color * gC).sum() + (alpha * gA).sum()
        loss.backward()

    def autograd_direct_step(i):
        for v in leaves.values():
            v.grad = None
        color, radii, depth, alpha = GaussianRasterizer(raster_settings=settings[i % 8])(means2D=m2d, **leaves)
        torch.autograd.backward([color, alpha], [gC, gA])

    def lean_step(i):
        vsr.render_views(params, [settings[i % 8]], [(gC, None, gA)])

    for name, fn in (("autograd + loss ops", autograd_step), ("autograd, upstream grads given", autograd_direct_step), ("render_views (lean)", lean_step)):
        for i in range(20): fn(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(a.steps): fn(i)
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        print("%-32s host enqueue %.1f us/step, wall incl. GPU drain %.1f us/step" % (name, t_host / a.steps * 1e6, t_all / a.steps * 1e6), flush=True)
        if a.profile:
            pr = cProfile.Profile(); pr.enable()
            for i in range(200): fn(i)
            pr.disable(); torch.cuda.synchronize()
            s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14)
            print("\n".join(l[:150] for l in s.getvalue().splitlines()[4:28]), flush=True)

if __name__ == "__main__":
    main()
