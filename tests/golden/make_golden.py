"""Generate tests/golden/reference_vectors.npz by IMPORTING the reference (read-only, /root/reference) in this container.

The rasterizer op itself is not vendored in the reference, so it cannot be run; what CAN be pinned from
the reference's own code are the pieces either side of it that define conventions the op must honour:
  * cam_utils.orbit_camera + gs_renderer.MiniCam / getProjectionMatrix  -> view / projection matrices, camera centre
  * sh_utils.eval_sh (+ RGB2SH / SH2RGB)                                -> SH basis, signs, constants
  * gs_renderer.build_rotation / build_scaling_rotation / strip_symmetric -> quaternion convention, cov3D packing
Missing third-party imports of gs_renderer.py (plyfile, kiui, simple_knn, diff_gaussian_rasterization, mesh...) are
stubbed; torch.Tensor.cuda / device="cuda" are redirected to the CPU.  Run:  python tests/golden/make_golden.py
"""
import math
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.npz")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    sys.path.insert(0, REF)
    _stub("plyfile", PlyData=object, PlyElement=object)
    _stub("kiui")
    _stub("mcubes")
    _stub("simple_knn")
    _stub("simple_knn._C", distCUDA2=None)
    _stub("diff_gaussian_rasterization", GaussianRasterizationSettings=object, GaussianRasterizer=object)
    _stub("mesh", Mesh=object)
    _stub("mesh_utils", decimate_mesh=None, clean_mesh=None)
    # the reference hard-codes device="cuda": run those helpers on the CPU
    torch.Tensor.cuda = lambda self, *a, **k: self
    _zeros, _tensor = torch.zeros, torch.tensor

    def zeros(*a, **k):
        k.pop("device", None)
        return _zeros(*a, **k)

    def tensor(*a, **k):
        k.pop("device", None)
        return _tensor(*a, **k)

    torch.zeros, torch.tensor = zeros, tensor
    import cam_utils
    import gs_renderer
    import sh_utils
    return cam_utils, gs_renderer, sh_utils


def main():
    cam_utils, gs_renderer, sh_utils = import_reference()
    rng = np.random.default_rng(20240922)
    out = {}
    # ---- cameras: (elevation, azimuth, radius, W, H)
    cams = [(0.0, 0.0, 2.0, 800, 800), (-30.0, 45.0, 2.0, 800, 800), (17.0, -130.0, 2.5, 512, 256), (29.0, 179.0, 1.7, 100, 70)]
    out["cam_params"] = np.array(cams, np.float64)
    V, PM, C, TAN = [], [], [], []
    for el, az, r, W, H in cams:
        oc = cam_utils.OrbitCamera(int(W), int(H), r=r, fovy=49.1)
        pose = cam_utils.orbit_camera(el, az, r)
        mc = gs_renderer.MiniCam(pose, int(W), int(H), oc.fovy, oc.fovx, oc.near, oc.far)
        V.append(mc.world_view_transform.numpy()); PM.append(mc.full_proj_transform.numpy()); C.append(mc.camera_center.numpy())
        TAN.append([math.tan(mc.FoVx * 0.5), math.tan(mc.FoVy * 0.5)])
    out["cam_view"], out["cam_fullproj"], out["cam_center"], out["cam_tanfov"] = map(np.array, (V, PM, C, TAN))
    # ---- SH evaluation, degrees 0..3 (sh laid out [..., C, coeffs] for eval_sh)
    n = 64
    dirs = rng.normal(size=(n, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    sh = rng.normal(size=(n, 16, 3))            # the op's layout [P, M, 3]
    out["sh_dirs"], out["sh_coeffs"] = dirs, sh
    for deg in range(4):
        res = sh_utils.eval_sh(deg, torch.tensor(sh).transpose(1, 2), torch.tensor(dirs))
        out["sh_rgb_deg%d" % deg] = res.numpy()
    out["rgb2sh_of_half_quarter"] = sh_utils.RGB2SH(np.array([0.5, 0.25, 1.0]))
    # ---- quaternion -> rotation (normalised in the reference), covariance packing
    q = rng.normal(size=(n, 4)); s = np.exp(rng.normal(-3.0, 0.5, size=(n, 3)))
    out["quat"], out["scale"] = q, s
    out["rotmat"] = gs_renderer.build_rotation(torch.tensor(q, dtype=torch.float32)).numpy()
    L = gs_renderer.build_scaling_rotation(torch.tensor(1.7 * s, dtype=torch.float32), torch.tensor(q, dtype=torch.float32))
    out["cov6_mod1p7"] = gs_renderer.strip_symmetric(L @ L.transpose(1, 2)).numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
