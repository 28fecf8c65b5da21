"""CPU tests of the oracle (test infrastructure) — the reference ships no tests or golden vectors for the rasterizer
path (PARITY UNPINNED), so the oracle is pinned by everything that CAN be pinned:
  * golden vectors generated from the reference's own Python (tests/golden/make_golden.py): camera matrices, SH
    evaluation, quaternion / covariance conventions;
  * the known-answer values of SURVEY.md §8(c) (analytic single Gaussian, camera matrices);
  * an independent derivation: the C oracle's hand-written backward against PyTorch autograd of the torch oracle, and the
    torch oracle against float64 finite differences;
  * invariants of the algorithm.
"""
import math
import os

import numpy as np
import pytest
import torch

import helpers as h
from dreamgaussian_b200 import scene
from oracle import c_oracle, torch_oracle

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


# ------------------------------------------------------------------ golden vectors from the reference's Python
def test_camera_matrices_match_reference_minicam():
    for k, (el, az, r, W, H) in enumerate(GOLD["cam_params"]):
        cam = scene.orbit_camera(el, az, r, int(W), int(H))
        np.testing.assert_allclose(cam.world_view_transform, GOLD["cam_view"][k], atol=2e-6)
        np.testing.assert_allclose(cam.full_proj_transform, GOLD["cam_fullproj"][k], atol=5e-6)
        np.testing.assert_allclose(cam.camera_center, GOLD["cam_center"][k], atol=1e-6)
        np.testing.assert_allclose([cam.tanfovx, cam.tanfovy], GOLD["cam_tanfov"][k], rtol=1e-6)


def test_survey_known_answer_camera_values():
    cam = scene.orbit_camera(0, 0, 2.0, 800, 800)
    assert abs(cam.tanfovx - 0.4567805803) < 1e-6 and abs(cam.fovy - 0.8569566627) < 1e-6
    np.testing.assert_allclose(cam.world_view_transform, [[1, 0, 0, 0], [0, -1, 0, 0], [0, 0, -1, 0], [0, 0, 2, 1]], atol=1e-6)
    np.testing.assert_allclose(cam.full_proj_transform,
                               [[2.189235, 0, 0, 0], [0, -2.189235, 0, 0], [0, 0, -1.0001, -1], [0, 0, 1.9901991, 2]], atol=2e-6)
    np.testing.assert_allclose(cam.camera_center, [0, 0, -2], atol=1e-7)
    cam2 = scene.orbit_camera(-30, 45, 2.0, 800, 800)
    np.testing.assert_allclose(cam2.camera_center, [-1.2247449, -1.0, -1.2247449], atol=1e-6)
    p = np.array([0.1, 0.2, 0.3, 1.0])
    np.testing.assert_allclose(p @ cam2.world_view_transform.astype(np.float64), [-0.1414214, -0.0317837, 1.655051, 1.0], atol=2e-6)
    np.testing.assert_allclose(p @ cam2.full_proj_transform.astype(np.float64), [-0.3096046, -0.069582, 1.6452155, 1.655051], atol=3e-6)


def _settings(cam, deg, bg=(0.0, 0.0, 0.0), mod=1.0):
    return dict(image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                bg=np.asarray(bg, np.float64), scale_modifier=mod, viewmatrix=cam.world_view_transform,
                projmatrix=cam.full_proj_transform, sh_degree=deg, campos=cam.camera_center)


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_sh_to_rgb_matches_reference_eval_sh(deg):
    cam = scene.orbit_camera(0, 0, 2.0, 256, 256)
    dirs, sh = GOLD["sh_dirs"], GOLD["sh_coeffs"]
    n = dirs.shape[0]
    means = cam.camera_center[None].astype(np.float64) + 0.5 * dirs      # normalize(p - campos) == dirs
    r = c_oracle.forward(**_settings(cam, deg), means3D=means, opacities=np.full((n, 1), 0.5), shs=sh,
                         scales=np.full((n, 3), 0.01), rotations=np.tile([1.0, 0, 0, 0], (n, 1)))
    assert (r.radii > 0).all()
    expect = np.maximum(GOLD["sh_rgb_deg%d" % deg] + 0.5, 0.0)          # gs_renderer.py:793
    np.testing.assert_allclose(r.state()["rgb"], expect, atol=2e-6)


def test_cov3d_from_scale_rotation_matches_reference_packing():
    """scales+rotations (scale_modifier 1.7) and the reference's packed get_covariance give the same conics."""
    cam = scene.orbit_camera(10, 20, 2.0, 128, 128)
    q = GOLD["quat"] / np.linalg.norm(GOLD["quat"], axis=1, keepdims=True)      # the caller normalises (gs_renderer.py:142)
    s = GOLD["scale"]
    n = q.shape[0]
    rng = np.random.default_rng(5)
    means = rng.uniform(-0.3, 0.3, (n, 3))
    common = dict(means3D=means, opacities=np.full((n, 1), 0.7), colors_precomp=rng.random((n, 3)))
    a = c_oracle.forward(**_settings(cam, 0, mod=1.7), scales=s, rotations=q, **common)
    b = c_oracle.forward(**_settings(cam, 0, mod=1.0), cov3D_precomp=GOLD["cov6_mod1p7"], **common)
    assert (a.radii == b.radii).all() and (a.radii > 0).any()
    np.testing.assert_allclose(a.state()["conic"], b.state()["conic"], rtol=2e-4, atol=1e-9)   # golden cov is float32
    np.testing.assert_allclose(a.color, b.color, atol=2e-5)
    assert abs(GOLD["rgb2sh_of_half_quarter"][0]) < 1e-12 and abs(scene.SH_C0 * GOLD["rgb2sh_of_half_quarter"][2] - 0.5) < 1e-12


# ------------------------------------------------------------------ analytic known answer
def test_single_isotropic_gaussian_known_answer():
    W = H = 800
    cam = scene.orbit_camera(0, 0, 2.0, W, H)
    sigma, o, rgb, bg = 0.02, 0.6, np.array([0.2, 0.5, 0.9]), np.array([1.0, 1.0, 1.0])
    r = c_oracle.forward(**_settings(cam, 0, bg=bg), means3D=np.zeros((1, 3)), opacities=[[o]], colors_precomp=rgb[None],
                         scales=np.full((1, 3), sigma), rotations=[[1.0, 0, 0, 0]])
    fx = W / (2 * cam.tanfovx)
    cov = (sigma * fx / 2.0) ** 2 + float(np.float32(0.3))
    st = r.state()
    assert abs(fx - 875.694) < 2e-3
    np.testing.assert_allclose([st["px"][0], st["py"][0]], [399.5, 399.5], atol=1e-4)
    np.testing.assert_allclose(st["conic"][0], [1 / cov, 0.0, 1 / cov], rtol=1e-6, atol=1e-12)
    assert r.radii[0] == math.ceil(3 * math.sqrt(cov))
    a = min(float(np.float32(0.99)), o * math.exp(-0.5 * (0.5 ** 2 + 0.5 ** 2) / cov))
    for (y, x) in ((399, 399), (400, 400), (399, 400)):
        np.testing.assert_allclose(r.color[:, y, x], a * rgb + (1 - a) * bg, atol=1e-9)
        assert abs(r.alpha[0, y, x] - a) < 1e-12 and abs(r.depth[0, y, x] - a * 2.0) < 1e-9
    # outside the 3-sigma tile rect nothing is drawn
    rad = r.radii[0]
    far = int(399.5 + rad + 17)
    assert r.alpha[0, 399, far] == 0.0 and np.allclose(r.color[:, 399, far], bg)


# ------------------------------------------------------------------ independent derivations
CASES = [dict(P=250, res=48, deg=3, sigma=0.05, elev=10, azim=30),
         dict(P=200, res=40, deg=1, sigma=0.2, elev=25, azim=-100),          # big Gaussians: J clamp + near cull
         dict(P=150, res=33, deg=0, sigma=0.06, elev=-20, azim=200, scale_modifier=1.3)]


@pytest.mark.parametrize("kw", CASES)
def test_c_oracle_backward_matches_autograd_of_torch_oracle(kw):
    s, i = h.make_case(**kw)
    H, W = s["image_height"], s["image_width"]
    gC, gD, gA = (g.astype(np.float64) for g in h.upstream_grads(H, W))
    r = c_oracle.forward(**s, **i, dtype=np.float64)
    gr = r.backward(gC, gD, gA)
    t64 = lambda a: torch.tensor(np.asarray(a, np.float64))
    tin = {k: t64(v).requires_grad_(True) for k, v in i.items()}
    m2d = torch.zeros(len(i["means3D"]), 3, dtype=torch.float64, requires_grad=True)
    ts = {k: (t64(v) if isinstance(v, np.ndarray) else v) for k, v in s.items()}
    c, rad, d, a = torch_oracle.rasterize(**ts, means2D=m2d, **tin)
    np.testing.assert_allclose(c.detach().numpy(), r.color, atol=1e-12)
    np.testing.assert_allclose(d.detach().numpy(), r.depth, atol=1e-12)
    np.testing.assert_allclose(a.detach().numpy(), r.alpha, atol=1e-12)
    assert (rad.numpy() == r.radii).all()
    ((c * t64(gC)).sum() + (d * t64(gD)).sum() + (a * t64(gA)).sum()).backward()
    for k in i:
        ref = tin[k].grad.numpy().reshape(gr[k].shape)
        np.testing.assert_allclose(gr[k], ref, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(ref).max()), err_msg=k)
    np.testing.assert_allclose(gr["means2D"], m2d.grad.numpy(), rtol=1e-9, atol=1e-9 * np.abs(m2d.grad.numpy()).max())


def test_torch_oracle_against_finite_differences():
    """float64 central differences on a tiny scene whose Gaussians stay away from every discrete decision."""
    s, i = h.make_case(P=12, res=24, deg=2, sigma=0.12, seed=3, elev=5, azim=15, opacity="trained")
    i["opacities"] = np.clip(i["opacities"], 0.2, 0.7)
    ref = c_oracle.forward(**s, **i, dtype=np.float64, eps=1e-3)
    apx, ag, _ = ref.flags()
    t64 = lambda a: torch.tensor(np.asarray(a, np.float64))
    ts = {k: (t64(v) if isinstance(v, np.ndarray) else v) for k, v in s.items()}
    rng = np.random.default_rng(0)
    gC = t64(rng.normal(size=(3, 24, 24))) * t64(1.0 - (apx > 0))[None]      # ignore pixels near a threshold
    gA = t64(rng.normal(size=(1, 24, 24))) * t64(1.0 - (apx > 0))[None]
    gD = t64(rng.normal(size=(1, 24, 24))) * t64(1.0 - (apx > 0))[None]

    def loss(inp):
        c, _, d, a = torch_oracle.rasterize(**ts, means2D=None, **inp)
        return (c * gC).sum() + (d * gD).sum() + (a * gA).sum()

    base = {k: t64(v).requires_grad_(True) for k, v in i.items()}
    L = loss(base)
    L.backward()
    for k in ("means3D", "scales", "rotations", "opacities", "shs"):
        g = base[k].grad
        for _ in range(3):
            dirn = t64(rng.normal(size=tuple(g.shape)))
            hstep = 1e-6
            pl = {kk: (vv.detach() + hstep * dirn if kk == k else vv.detach()) for kk, vv in base.items()}
            mi = {kk: (vv.detach() - hstep * dirn if kk == k else vv.detach()) for kk, vv in base.items()}
            fd = (loss(pl) - loss(mi)).item() / (2 * hstep)
            an = (g * dirn).sum().item()
            assert abs(fd - an) <= 2e-5 * max(1.0, abs(an)), (k, fd, an)


# ------------------------------------------------------------------ invariants
def test_invariants_alpha_culling_permutation():
    s, i = h.make_case(P=400, res=64, deg=1, sigma=0.05, elev=12, azim=70)
    # push a few Gaussians behind the camera / off screen
    i["means3D"] = i["means3D"].copy()
    i["means3D"][:5] = [0.0, 0.0, 5.0]          # behind the eye for this view
    i["means3D"][5:10] = [40.0, 0.0, 0.0]       # far off screen
    g = h.upstream_grads(64, 64)
    ref = h.run_oracle(s, i, g)
    # alpha = 1 - prod(1 - a) <= 1, colour - T*bg >= 0
    assert ref["alpha"].min() >= 0 and ref["alpha"].max() <= 1.0
    culled = ref["radii"] == 0
    assert culled[:10].all() and not culled[10:].all()
    for k, v in ref["grads"].items():
        if v.shape[0] == len(culled) and v.size:
            assert np.all(v[culled] == 0), k
    # permuting the Gaussians permutes the gradients and leaves the image unchanged (distinct depths)
    perm = np.random.default_rng(1).permutation(400)
    ip = {k: v[perm] for k, v in i.items()}
    refp = h.run_oracle(s, ip, g)
    np.testing.assert_allclose(refp["color"], ref["color"], atol=1e-12)
    np.testing.assert_allclose(refp["grads"]["means3D"], ref["grads"]["means3D"][perm], atol=1e-9)
    assert (refp["radii"] == ref["radii"][perm]).all()


def test_padded_sh_storage_is_ignored_beyond_the_active_degree():
    """gs_renderer.py:806: shs = get_features with (max_sh_degree+1)^2 rows while sh_degree = active_sh_degree."""
    s, i = h.make_case(P=300, res=48, deg=1, sigma=0.05, elev=15, azim=70)
    pad = dict(i)
    pad["shs"] = np.concatenate([i["shs"], np.random.default_rng(3).normal(size=(300, 12, 3)).astype(np.float32)], axis=1)
    g = h.upstream_grads(48, 48)
    a, b = h.run_oracle(s, i, g), h.run_oracle(s, pad, g)
    assert np.array_equal(a["color"], b["color"])
    assert not b["grads"]["shs"][:, 4:].any()
    np.testing.assert_allclose(b["grads"]["shs"][:, :4], a["grads"]["shs"], rtol=0, atol=1e-11)    # OpenMP summation order


def test_raw_parameter_oracle_is_the_activation_chain_rule_of_the_plain_one():
    """helpers.run_oracle_raw (checker of the fused-activation path) against closed forms of the
    activations' derivatives (gs_renderer.py:127-138): exp' = exp, sigmoid' = s(1-s), normalize' = (I - q q^T)/|raw|."""
    s, i = h.make_case(P=200, res=40, deg=2, sigma=0.06, elev=-5, azim=33)
    raw = scene.to_raw_parameters(i)
    g = h.upstream_grads(40, 40)
    plain, fused = h.run_oracle(s, i, g), h.run_oracle_raw(s, raw, g)
    np.testing.assert_allclose(fused["color"], plain["color"], atol=2e-6)          # float32 round trip of log / logit
    gp, gf = plain["grads"], fused["grads"]
    sc, op = np.exp(raw["scaling"].astype(np.float64)), 1 / (1 + np.exp(-raw["opacity"].astype(np.float64)))
    rtol = 2e-4                                                                    # plain oracle ran on the float32 activated values
    np.testing.assert_allclose(gf["scaling"], gp["scales"] * sc, rtol=rtol, atol=1e-6 * np.abs(gp["scales"]).max())
    np.testing.assert_allclose(gf["opacity"], gp["opacities"] * op * (1 - op), rtol=rtol, atol=1e-6 * np.abs(gp["opacities"]).max())
    np.testing.assert_allclose(gf["features_dc"], gp["shs"][:, :1], rtol=rtol, atol=1e-6)
    np.testing.assert_allclose(gf["features_rest"], gp["shs"][:, 1:], rtol=rtol, atol=1e-6)
    r = raw["rotation"].astype(np.float64)
    n = np.linalg.norm(r, axis=1, keepdims=True)
    q = r / n
    want = (gp["rotations"] - q * (q * gp["rotations"]).sum(1, keepdims=True)) / n
    np.testing.assert_allclose(gf["rotation"], want, rtol=rtol, atol=2e-6 * np.abs(want).max())


def test_oracle_edge_cases_empty_and_validation():
    cam = scene.orbit_camera(0, 0, 2.0, 32, 24)
    st = _settings(cam, 0, bg=(0.1, 0.2, 0.3))
    r = c_oracle.forward(**st, means3D=np.zeros((0, 3)), opacities=np.zeros((0, 1)), colors_precomp=np.zeros((0, 3)),
                         scales=np.zeros((0, 3)), rotations=np.zeros((0, 4)))
    assert r.color.shape == (3, 24, 32) and np.allclose(r.color, np.array([0.1, 0.2, 0.3])[:, None, None]) and r.alpha.max() == 0
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        c_oracle.forward(**st, means3D=np.zeros((1, 3)), opacities=np.zeros((1, 1)), scales=np.ones((1, 3)), rotations=np.ones((1, 4)))
    with pytest.raises(Exception, match="scale/rotation pair"):
        c_oracle.forward(**st, means3D=np.zeros((1, 3)), opacities=np.zeros((1, 1)), colors_precomp=np.zeros((1, 3)))
    # float32 build == float64 build up to float32 rounding away from decision boundaries
    s, i = h.make_case(P=300, res=48, deg=2, sigma=0.05)
    a, b = h.run_oracle(s, i), h.run_oracle(s, i, dtype=np.float32)
    m = ~a["ambig_px"].astype(bool)
    assert np.abs(a["color"] - b["color"])[:, m].max() < 2e-5
