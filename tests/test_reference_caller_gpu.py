"""The reference's OWN caller, unmodified, on the drop-in packages of this repo.

`gs_renderer.Renderer.initialize` -> `GaussianModel.create_from_pcd` -> `simple_knn._C.distCUDA2`
(/root/reference/gs_renderer.py:689-715, 331-354) and `Renderer.render` (/root/reference/gs_renderer.py:717-822: transposed-view
`viewmatrix` of MiniCam :662, the retain_grad dummy `means2D` :727-739, `clamp(0, 1)` :811) run as they are — from the
sources where /root/reference exists, else from their byte code in oracle/_ref/pyc (oracle/ref_caller.py) — and import
`diff_gaussian_rasterization` / `simple_knn._C` from THIS repo.  Everything the caller gets back (image, alpha, depth, radii,
leaf gradients, viewspace-point gradients, densification statistics) is compared with the CPU oracle fed the same tensors.
"""
import types

import numpy as np
import pytest
import torch

import helpers as h
from oracle import ref_caller

pytestmark = pytest.mark.gpu


def _reference():
    if not ref_caller.available():
        pytest.skip("neither /root/reference nor oracle/_ref/pyc is present")
    return ref_caller.load()


def _training_args():
    # /root/reference/configs/image.yaml:67-75
    return types.SimpleNamespace(position_lr_init=0.001, position_lr_final=0.00002, position_lr_delay_mult=0.02,
                                 position_lr_max_steps=500, feature_lr=0.01, opacity_lr=0.05, scaling_lr=0.005,
                                 rotation_lr=0.005, percent_dense=0.01)


CASES = {
    # BASELINE.json configs[0]: what `python main.py --config configs/image.yaml` renders first (image.yaml:65-66, 12, 41-44)
    "image_yaml_init": dict(sh_degree=0, num_pts=5000, W=256, H=256, elev=0.0, azim=0.0, scaling_modifier=1.0, bg=None, ups=0),
    # SH degree 3 active, non-square image, scaling modifier, explicit background, perturbed parameters
    "deg3_perturbed": dict(sh_degree=3, num_pts=20000, W=320, H=200, elev=-20.0, azim=130.0, scaling_modifier=1.15,
                           bg=(0.1, 0.4, 0.7), ups=3),
}


@pytest.mark.parametrize("name", list(CASES))
def test_reference_renderer_runs_unmodified_on_the_drop_in(name):
    c = CASES[name]
    cam_utils, gs, _sh = _reference()
    import diff_gaussian_rasterization as ours
    assert gs.GaussianRasterizer is ours.GaussianRasterizer and gs.GaussianRasterizationSettings is ours.GaussianRasterizationSettings
    import simple_knn._C as knn
    assert gs.distCUDA2 is knn.distCUDA2

    np.random.seed(7)
    torch.manual_seed(7)
    r = gs.Renderer(sh_degree=c["sh_degree"])
    r.initialize(num_pts=c["num_pts"])                       # random ball -> create_from_pcd -> distCUDA2 (this repo's kernels)
    gm = r.gaussians
    gm.training_setup(_training_args())
    for _ in range(c["ups"]):
        gm.oneupSHdegree()
    if c["ups"]:
        with torch.no_grad():                                # leave the symmetric initial state: anisotropy, rotations, colour detail
            gm._features_rest.normal_(0.0, 0.1)
            gm._scaling.add_(0.3 * torch.randn_like(gm._scaling))
            gm._rotation.copy_(torch.randn_like(gm._rotation))
            gm._opacity.copy_(torch.logit(torch.rand_like(gm._opacity) * 0.9 + 0.05))
    oc = cam_utils.OrbitCamera(c["W"], c["H"], r=2, fovy=49.1)
    cam = gs.MiniCam(cam_utils.orbit_camera(c["elev"], c["azim"], 2), c["W"], c["H"], oc.fovy, oc.fovx, oc.near, oc.far)
    bg = None if c["bg"] is None else torch.tensor(c["bg"], dtype=torch.float32, device="cuda")
    out = r.render(cam, scaling_modifier=c["scaling_modifier"], bg_color=bg)
    H, W = c["H"], c["W"]
    assert out["image"].shape == (3, H, W) and out["alpha"].shape == (1, H, W) and out["depth"].shape == (1, H, W)
    assert out["radii"].dtype == torch.int32 and out["visibility_filter"].dtype == torch.bool

    gC, gD, gA = h.upstream_grads(H, W, seed=5, depth=False)
    t = lambda a: torch.tensor(a, device="cuda")
    loss = (out["image"] * t(gC)).sum() + (out["alpha"] * t(gA)).sum()
    loss.backward()
    gm.add_densification_stats(out["viewspace_points"], out["visibility_filter"])           # gs_renderer.py:625-627

    # ---- the oracle on the very tensors the caller handed to the op
    f64 = lambda x: x.detach().double().cpu().numpy()
    bg_used = r.bg_color if bg is None else bg
    settings = dict(image_height=H, image_width=W, tanfovx=float(np.tan(cam.FoVx * 0.5)), tanfovy=float(np.tan(cam.FoVy * 0.5)),
                    bg=f64(bg_used), scale_modifier=c["scaling_modifier"], viewmatrix=f64(cam.world_view_transform),
                    projmatrix=f64(cam.full_proj_transform), sh_degree=gm.active_sh_degree, campos=f64(cam.camera_center))
    raw = dict(xyz=f64(gm._xyz), features_dc=f64(gm._features_dc), features_rest=f64(gm._features_rest), opacity=f64(gm._opacity),
               scaling=f64(gm._scaling), rotation=f64(gm._rotation))
    # clamp(0, 1) of the caller: gradient passes where the un-clamped colour lies in [0, 1] (torch's rule); the un-clamped
    # colour is what one more (deterministic) call of the op with the same tensors returns
    with torch.no_grad():
        rs = ours.GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=settings["tanfovx"], tanfovy=settings["tanfovy"], bg=bg_used,
            scale_modifier=c["scaling_modifier"], viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
            sh_degree=gm.active_sh_degree, campos=cam.camera_center, prefiltered=False, debug=False)
        pre, radii2, _, _ = ours.GaussianRasterizer(rs)(means3D=gm.get_xyz, means2D=torch.zeros_like(gm.get_xyz), shs=gm.get_features,
                                                        opacities=gm.get_opacity, scales=gm.get_scaling, rotations=gm.get_rotation)
    assert torch.equal(radii2, out["radii"]) and torch.equal(pre.clamp(0, 1), out["image"])
    passes = ((pre >= 0) & (pre <= 1)).cpu().numpy()
    ref = h.run_oracle_raw(settings, raw, (gC * passes, None, gA))
    cu = dict(color=pre.cpu().numpy(), depth=out["depth"].detach().cpu().numpy(), alpha=out["alpha"].detach().cpu().numpy(),
              radii=out["radii"].cpu().numpy(),
              grads=dict(xyz=gm._xyz.grad, features_dc=gm._features_dc.grad, features_rest=gm._features_rest.grad,
                         opacity=gm._opacity.grad, scaling=gm._scaling.grad, rotation=gm._rotation.grad,
                         means2D=out["viewspace_points"].grad))
    cu["grads"] = {k: (v.detach().cpu().numpy() if v is not None else np.zeros(ref["grads"][k].shape, np.float32))
                   for k, v in cu["grads"].items()}
    ok, rep = h.compare(cu, ref, max_ambig_frac=0.10)
    assert ok, rep
    assert int((out["radii"] > 0).sum()) > c["num_pts"] // 2

    # densification statistics as the reference's own method computed them from OUR viewspace gradient
    vis = ref["radii"] > 0
    want = np.linalg.norm(ref["grads"]["means2D"][:, :2], axis=-1)
    got = gm.xyz_gradient_accum[:, 0].cpu().numpy()
    clean = (ref["ambig_g"] == 0) & vis
    assert np.abs(got[clean] - want[clean]).max() <= 1e-3 * max(want.max(), 1e-12)
    assert np.array_equal(gm.denom[:, 0].cpu().numpy()[ref["ambig_g"] == 0], vis[ref["ambig_g"] == 0].astype(np.float32))
