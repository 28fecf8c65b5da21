"""CPU restatement of `GaussianModel.extract_fields` — TEST INFRASTRUCTURE ONLY.

Reference: /root/reference/gs_renderer.py:218-294 (extract_fields), :64-83 (gaussian_3d_coeff), :85-117
(build_rotation / build_scaling_rotation), :128-132 (covariance = L L^T, strip_symmetric).  PARITY PINNED:
tests/golden/extract_fields_vectors.npz holds outputs of the reference's own method run in this container
(tests/golden/make_golden_fields.py); tests/test_fields_oracle.py checks this file against them.

Definition restated: keep Gaussians with sigmoid(opacity) > 0.005 (:229); normalise positions to ~[-1,1] with
center = (min+max)/2 and scale = 1.8 / max extent (:236-241); a resolution^3 grid over linspace(-1,1) is cut into
num_blocks^3 blocks; a voxel of block B sums  opacity * exp(-1/2 d^T Sigma^-1 d)  over the Gaussians whose centre lies
strictly inside B's voxel bounding box grown by relax_ratio * (2 / num_blocks) on every side (:262-266) — the truncation
is per BLOCK, not per distance, and is part of the result; exponents > 0 are treated as -1e10 (:81)."""
import numpy as np


def linspace(res, dtype):
    """torch.linspace(-1, 1, res): symmetric evaluation from both ends, each value one fused multiply-add of the float32
    step (emulated exactly: float32 x small integer and the sum are exact in float64, so there is a single rounding)."""
    i = np.arange(res)
    if dtype == np.float64:
        step = 2.0 / (res - 1)
        return np.where(i < res // 2, -1.0 + step * i, 1.0 - step * (res - 1 - i))
    step = np.float64(np.float32(2.0) / np.float32(res - 1))
    return np.where(i < res // 2, -1.0 + step * i, 1.0 - step * (res - 1 - i)).astype(np.float32)


def prepare(xyz, opacity_raw, scaling_raw, rotation_raw, dtype=np.float32):
    """Per-Gaussian state: mask, normalised centre, opacity, inverse covariance (inv_a .. inv_f), center, scale."""
    f = dtype
    op = (1.0 / (1.0 + np.exp(-opacity_raw.astype(f)))).astype(f).reshape(-1)
    mask = op > f(0.005)                                                        # :229
    x = xyz.astype(f)[mask]
    s = np.exp(scaling_raw.astype(f))[mask]
    mn, mx = x.min(axis=0), x.max(axis=0)
    center = ((mn + mx) / f(2)).astype(f)                                       # :237
    scale = 1.8 / float((mx - mn).max())                                        # :238  (python float)
    x = ((x - center) * f(scale)).astype(f)
    s = (s * f(scale)).astype(f)
    r = rotation_raw.astype(f)[mask]
    q = r / np.sqrt((r * r).sum(axis=1, keepdims=True))                         # :86-88
    w, qx, qy, qz = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.stack([1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - w * qz), 2 * (qx * qz + w * qy),
                  2 * (qx * qy + w * qz), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - w * qx),
                  2 * (qx * qz - w * qy), 2 * (qy * qz + w * qx), 1 - 2 * (qx * qx + qy * qy)], axis=1).reshape(-1, 3, 3).astype(f)
    L = R * s[:, None, :]                                                       # R @ diag(s)
    S = L @ L.transpose(0, 2, 1)
    a, b, c, d, e, g = S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]
    inv_det = f(1) / (a * d * g + 2 * e * c * b - e ** 2 * a - c ** 2 * d - b ** 2 * g + f(1e-24))      # :71
    inv = np.stack([(d * g - e ** 2) * inv_det, (e * c - b * g) * inv_det, (e * b - c * d) * inv_det,
                    (a * g - c ** 2) * inv_det, (b * c - e * a) * inv_det, (a * d - b ** 2) * inv_det], axis=1).astype(f)
    return dict(mask=mask, x=x, opacity=op[mask], inv=inv, center=center, scale=scale)


def extract_fields(xyz, opacity_raw, scaling_raw, rotation_raw, resolution=128, num_blocks=16, relax_ratio=1.5, dtype=np.float32):
    f = dtype
    st = prepare(xyz, opacity_raw, scaling_raw, rotation_raw, dtype)
    if resolution % num_blocks:
        raise ValueError("resolution must be a multiple of num_blocks")
    split = resolution // num_blocks
    block_size = f(2.0 / num_blocks)
    X = linspace(resolution, f)
    occ = np.zeros((resolution,) * 3, f)
    x, opa, inv = st["x"], st["opacity"], st["inv"]
    grow = f(block_size * f(relax_ratio))
    for xi in range(num_blocks):
        xs = X[xi * split:(xi + 1) * split]
        mx_ = (x[:, 0] < xs[-1] + grow) & (x[:, 0] > xs[0] - grow)
        if not mx_.any():
            continue
        for yi in range(num_blocks):
            ys = X[yi * split:(yi + 1) * split]
            my_ = mx_ & (x[:, 1] < ys[-1] + grow) & (x[:, 1] > ys[0] - grow)
            if not my_.any():
                continue
            for zi in range(num_blocks):
                zs = X[zi * split:(zi + 1) * split]
                m = my_ & (x[:, 2] < zs[-1] + grow) & (x[:, 2] > zs[0] - grow)          # :262-266, strict
                if not m.any():
                    continue
                gx, o, iv = x[m], opa[m], inv[m]
                dx = xs[:, None, None, None] - gx[None, None, None, :, 0]
                dy = ys[None, :, None, None] - gx[None, None, None, :, 1]
                dz = zs[None, None, :, None] - gx[None, None, None, :, 2]
                power = f(-0.5) * (dx * dx * iv[:, 0] + dy * dy * iv[:, 3] + dz * dz * iv[:, 5]) - dx * dy * iv[:, 1] - dx * dz * iv[:, 2] - dy * dz * iv[:, 4]
                power = np.where(power > 0, f(-1e10), power)                                # :81
                occ[xi * split:(xi + 1) * split, yi * split:(yi + 1) * split, zi * split:(zi + 1) * split] = (o * np.exp(power)).sum(axis=-1)
    return occ, st["center"], st["scale"]
