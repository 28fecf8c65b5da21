"""Golden vectors for SURVEY.md §8 row f4 (`GaussianModel.extract_fields`, /root/reference/gs_renderer.py:218-294), made by
RUNNING the reference's own method in this container on the CPU (its third-party imports stubbed, device="cuda"
redirected — see make_golden.py).  Writes tests/golden/extract_fields_vectors.npz.  Run: python tests/golden/make_golden_fields.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "extract_fields_vectors.npz")


def cloud(P, seed, spread=0.45, log_sigma=-3.6):
    rng = np.random.default_rng(seed)
    xyz = rng.normal(size=(P, 3)) * spread * np.array([1.0, 0.7, 0.5]) + np.array([0.3, -0.2, 0.1])
    return dict(xyz=xyz.astype(np.float32), opacity=rng.normal(0.0, 2.5, size=(P, 1)).astype(np.float32),      # ~2% fall below 0.005
                scaling=(log_sigma + rng.normal(0, 0.4, size=(P, 3))).astype(np.float32),
                rotation=rng.normal(size=(P, 4)).astype(np.float32))


def main():
    _, gs_renderer, _ = make_golden.import_reference()
    sys.modules["kiui"].lo = lambda *a, **k: None
    out = {}
    for name, P, res, nb, relax, seed in (("a", 1500, 32, 16, 1.5, 1), ("b", 4000, 64, 16, 1.5, 2), ("c", 800, 48, 8, 1.0, 3)):
        c = cloud(P, seed)
        gm = gs_renderer.GaussianModel(0)
        gm._xyz, gm._opacity = torch.tensor(c["xyz"]), torch.tensor(c["opacity"])
        gm._scaling, gm._rotation = torch.tensor(c["scaling"]), torch.tensor(c["rotation"])
        occ = gm.extract_fields(resolution=res, num_blocks=nb, relax_ratio=relax)
        for k, v in c.items():
            out["%s_%s" % (name, k)] = v
        out[name + "_params"] = np.array([res, nb, relax], np.float64)
        out[name + "_occ"] = occ.numpy()
        out[name + "_center"] = gm.center.numpy()
        out[name + "_scale"] = np.array(gm.scale, np.float64)
        print(name, P, res, nb, "occ max %.4f mean %.5f nonzero %.3f" % (occ.max(), occ.mean(), (occ > 0).float().mean()))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) >> 10, "KiB")


if __name__ == "__main__":
    main()
