"""Diagnostic: forward determinism and backward linearity in the upstream gradient at several sizes."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from dreamgaussian_b200 import scene, rasterizer as R
dev = torch.device("cuda")
import os as _os
CFG = eval(_os.environ.get("LIN_CFG", "((100000, 800, None), (300000, 1200, 0.009), (1000000, 1600, 0.006), (1000000, 800, 0.006))"))
for (P, res, sigma) in CFG:
    deg = 3
    cloud = scene.make_cloud(P, deg, seed=2, sigma=sigma)
    params = {k: torch.tensor(v, device=dev) for k, v in cloud.items()}
    cam = scene.orbit_camera(10, 30, 2.0, res, res)
    t = lambda a: torch.tensor(np.asarray(a, np.float32), device=dev)
    rs = R.GaussianRasterizationSettings(image_height=res, image_width=res, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=t(np.zeros(3)),
        scale_modifier=1.0, viewmatrix=t(cam.world_view_transform), projmatrix=t(cam.full_proj_transform), sh_degree=deg,
        campos=t(cam.camera_center), prefiltered=False, debug=False)
    torch.manual_seed(0)
    g1 = torch.randn(3, res, res, device=dev); g2 = torch.randn(3, res, res, device=dev)
    def grads(gc):
        pin = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        c, r, d, a = R.GaussianRasterizer(rs)(means3D=pin["means3D"], means2D=torch.zeros_like(pin["means3D"]), opacities=pin["opacities"],
                                              shs=pin["shs"], scales=pin["scales"], rotations=pin["rotations"])
        (c * gc).sum().backward()
        return c.detach().clone(), {k: v.grad.clone() for k, v in pin.items()}
    c1, ga = grads(g1); c1b, ga2 = grads(g1); c2, gb = grads(g2); c3, gab = grads(g1 + g2)
    key = (torch.cuda.current_device(), P, res, res)
    print(P, res, "hint", R._CAPACITY_HINT.get(key), "fwd identical:", torch.equal(c1, c1b), torch.equal(c1, c2), torch.equal(c1, c3), flush=True)
    for k in ga:
        sc = float(gab[k].abs().max())
        rep = float((ga[k] - ga2[k]).abs().max()) / sc
        lin = float((ga[k] + gb[k] - gab[k]).abs().max()) / sc
        idx = int((ga[k] + gb[k] - gab[k]).abs().reshape(P, -1).max(1).values.argmax())
        print("   %-10s scale %.3e repeat-diff %.2e linearity-err %.2e worst g=%d" % (k, sc, rep, lin, idx), flush=True)
