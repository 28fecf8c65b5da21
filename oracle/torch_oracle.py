"""Pure-PyTorch (CPU, autograd) restatement of the rasterizer op — the "CPU torch fallback" of BASELINE.json.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/dgr_oracle.c).  Its job is to be an INDEPENDENT derivation
of the backward: the forward is written with differentiable torch ops and gradients come from autograd, so the
hand-derived chain rule in oracle/dgr_oracle.c (and in the CUDA kernels) can be checked against it, and it can be
checked against finite differences in float64.

Two deliberate deviations from plain autograd mirror the op's published backward:
  * alpha = min(0.99, o*G) passes its gradient unmasked (straight-through);
  * inside the Jacobian J the clamped tx' = clamp(tx/tz)*tz is treated as the independent variable tx (gradient 1
    w.r.t. t.x when not clamped, 0 when clamped, and no extra t.z dependence).
Everything else (SH clamp mask, cull decisions constant) is what torch.clamp / masking give.

Reference anchors: call signature /root/reference/gs_renderer.py:745-809; SH basis sh_utils.py:57-100;
rotation/covariance gs_renderer.py:85-132; camera conventions gs_renderer.py:629-671.
"""
import math

import torch

import numpy as _np


def _f(v):
    """Constants are the float32-rounded values the op (and include/dgr_constants.h) uses, held as Python floats."""
    return float(_np.float32(v))


TILE = 16
NEAR_CULL_Z, W_EPS, LOWPASS, EIG_FLOOR, RADIUS_SIGMAS = _f(0.2), _f(1e-7), _f(0.3), _f(0.1), 3.0
FOV_CLAMP, ALPHA_MAX, ALPHA_MIN, T_STOP, SH_OFFSET = _f(1.3), _f(0.99), _f(_np.float32(1.0) / _np.float32(255.0)), _f(1e-4), 0.5
C0 = _f(0.28209479177387814)
C1 = _f(0.4886025119029199)
C2 = [_f(v) for v in (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)]
C3 = [_f(v) for v in (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
                      -0.4570457994644658, 1.445305721320277, -0.5900435899266435)]


def _sh_rgb(deg, sh, d):
    """sh [P,M,3], d [P,3] unit -> [P,3]."""
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    res = C0 * sh[:, 0]
    if deg > 0:
        res = res - C1 * y * sh[:, 1] + C1 * z * sh[:, 2] - C1 * x * sh[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = (res + C2[0] * xy * sh[:, 4] + C2[1] * yz * sh[:, 5] + C2[2] * (2 * zz - xx - yy) * sh[:, 6]
               + C2[3] * xz * sh[:, 7] + C2[4] * (xx - yy) * sh[:, 8])
    if deg > 2:
        res = (res + C3[0] * y * (3 * xx - yy) * sh[:, 9] + C3[1] * xy * z * sh[:, 10]
               + C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
               + C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + C3[5] * z * (xx - yy) * sh[:, 14]
               + C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return res


def _quat_to_R(q):
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).reshape(-1, 3, 3)


class _StraightThroughMin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, hi):
        return torch.clamp(x, max=hi)

    @staticmethod
    def backward(ctx, g):
        return g, None


def rasterize(*, image_height, image_width, tanfovx, tanfovy, bg, scale_modifier, viewmatrix, projmatrix, sh_degree,
              campos, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
              cov3D_precomp=None):
    """Returns (color[3,H,W], radii[P] int32, depth[1,H,W], alpha[1,H,W]); differentiable w.r.t. the float inputs.

    means2D receives the screen-space gradient in NDC units, as the op does (it enters as ndc + means2D[:, :2]).
    """
    if (shs is None) == (colors_precomp is None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
    H, W = int(image_height), int(image_width)
    dt = means3D.dtype
    P = means3D.shape[0]
    V, PM = viewmatrix.to(dt), projmatrix.to(dt)
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    fx, fy = W / (2 * tanfovx), H / (2 * tanfovy)
    ones = torch.ones(P, 1, dtype=dt)
    t = torch.cat([means3D, ones], 1) @ V[:, :3]
    ph = torch.cat([means3D, ones], 1) @ PM
    ndc = ph[:, :2] / (ph[:, 3:4] + W_EPS)
    if means2D is not None:
        ndc = ndc + means2D[:, :2]
    if cov3D_precomp is None:
        R = _quat_to_R(rotations)
        Mx = R * (scale_modifier * scales)[:, None, :]
        Sigma = Mx @ Mx.transpose(1, 2)
    else:
        c = cov3D_precomp
        Sigma = torch.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4], c[:, 2], c[:, 4], c[:, 5]], 1).reshape(-1, 3, 3)
    tz = t[:, 2]
    tz_safe = torch.where(tz > NEAR_CULL_Z, tz, torch.ones_like(tz))
    limx, limy = FOV_CLAMP * tanfovx, FOV_CLAMP * tanfovy
    txtz, tytz = t[:, 0] / tz_safe, t[:, 1] / tz_safe
    clx, cly = (txtz < -limx) | (txtz > limx), (tytz < -limy) | (tytz > limy)
    tx = torch.where(clx, (txtz.clamp(-limx, limx) * tz_safe).detach(), t[:, 0])
    ty = torch.where(cly, (tytz.clamp(-limy, limy) * tz_safe).detach(), t[:, 1])
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz_safe, zero, -fx * tx / (tz_safe * tz_safe),
                     zero, fy / tz_safe, -fy * ty / (tz_safe * tz_safe)], 1).reshape(-1, 2, 3)
    Rwv = V[:3, :3].t()
    Tm = J @ Rwv
    cov = Tm @ Sigma @ Tm.transpose(1, 2)
    cxx, cxy, cyy = cov[:, 0, 0] + LOWPASS, cov[:, 0, 1], cov[:, 1, 1] + LOWPASS
    det = cxx * cyy - cxy * cxy
    det_safe = torch.where(det != 0, det, torch.ones_like(det))
    conA, conB, conC = cyy / det_safe, -cxy / det_safe, cxx / det_safe
    mid = 0.5 * (cxx + cyy)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=EIG_FLOOR))
    radius = torch.ceil(RADIUS_SIGMAS * torch.sqrt(lam)).detach()
    mx = ((ndc[:, 0] + 1) * W - 1) * 0.5
    my = ((ndc[:, 1] + 1) * H - 1) * 0.5
    with torch.no_grad():
        trunc = lambda v, lim: torch.clamp(torch.trunc(v), 0, lim).to(torch.int64)
        rminx, rminy = trunc((mx - radius) / TILE, gx), trunc((my - radius) / TILE, gy)
        rmaxx, rmaxy = trunc((mx + radius + TILE - 1) / TILE, gx), trunc((my + radius + TILE - 1) / TILE, gy)
        visible = (tz > NEAR_CULL_Z) & (det != 0) & (rmaxx > rminx) & (rmaxy > rminy)
    if shs is not None:
        d = means3D - campos.to(dt)[None, :]
        d = d / torch.sqrt((d * d).sum(1, keepdim=True))
        rgb = torch.clamp_min(_sh_rgb(sh_degree, shs, d) + SH_OFFSET, 0.0)
    else:
        rgb = colors_precomp
    radii = torch.where(visible, radius, torch.zeros_like(radius)).to(torch.int32)
    opac = opacities.reshape(-1)

    color = torch.zeros(3, H, W, dtype=dt) + bg.to(dt)[:, None, None]
    depth_img = torch.zeros(1, H, W, dtype=dt)
    alpha_img = torch.zeros(1, H, W, dtype=dt)
    vis_idx = torch.nonzero(visible).reshape(-1)
    depth_key = tz.detach()
    rows_c, rows_d, rows_a = {}, {}, {}
    for ty_ in range(gy):
        for tx_ in range(gx):
            m = (rminx[vis_idx] <= tx_) & (rmaxx[vis_idx] > tx_) & (rminy[vis_idx] <= ty_) & (rmaxy[vis_idx] > ty_)
            ids = vis_idx[m]
            x0, y0 = tx_ * TILE, ty_ * TILE
            x1, y1 = min(x0 + TILE, W), min(y0 + TILE, H)
            if ids.numel() == 0:
                continue
            # stable (depth, index) order: ids are already index-ascending, stable argsort keeps ties
            order = torch.sort(depth_key[ids], stable=True).indices
            ids = ids[order]
            ys, xs = torch.meshgrid(torch.arange(y0, y1, dtype=dt), torch.arange(x0, x1, dtype=dt), indexing="ij")
            xs, ys = xs.reshape(-1, 1), ys.reshape(-1, 1)                       # [Npx,1]
            dx, dy = mx[ids][None, :] - xs, my[ids][None, :] - ys               # [Npx,K]
            power = -0.5 * (conA[ids][None] * dx * dx + conC[ids][None] * dy * dy) - conB[ids][None] * dx * dy
            G = torch.exp(torch.clamp(power, max=0.0))
            a = _StraightThroughMin.apply(opac[ids][None] * G, ALPHA_MAX)
            with torch.no_grad():
                contrib = (power <= 0) & (a >= ALPHA_MIN)
            a = torch.where(contrib, a, torch.zeros_like(a))
            one_m = 1 - a
            Texcl = torch.cumprod(torch.cat([torch.ones_like(one_m[:, :1]), one_m[:, :-1]], 1), 1)
            with torch.no_grad():
                stop = contrib & ((Texcl * one_m).detach() < T_STOP)
                stopped = torch.cummax(stop.to(torch.int8), 1).values.bool()
            w = torch.where(stopped, torch.zeros_like(a), a * Texcl)           # [Npx,K]
            Tfin = 1 - w.sum(1)                                                 # == prod(1-a) over processed set
            # exact product form for the background term (keeps autograd identical to the op's recurrence)
            Tfin = torch.prod(torch.where(stopped, torch.ones_like(one_m), one_m), 1)
            c_t = w @ rgb[ids]                                                  # [Npx,3]
            d_t = w @ tz[ids]
            a_t = w.sum(1)
            hh, ww = y1 - y0, x1 - x0
            color[:, y0:y1, x0:x1] = (c_t + Tfin[:, None] * bg.to(dt)[None, :]).t().reshape(3, hh, ww)
            depth_img[0, y0:y1, x0:x1] = d_t.reshape(hh, ww)
            alpha_img[0, y0:y1, x0:x1] = a_t.reshape(hh, ww)
    return color, radii, depth_img, alpha_img
