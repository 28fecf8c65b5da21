"""PLY checkpoint format (gs_renderer.py:384-462): layout, header text and round trip.  CPU only."""
import numpy as np
import torch

import helpers  # noqa: F401
from dreamgaussian_b200 import ply


def _model(P=37, deg=2, seed=0):
    rng = np.random.default_rng(seed)
    M = (deg + 1) ** 2
    f = lambda *s: torch.tensor(rng.normal(size=s).astype(np.float32))
    return dict(xyz=f(P, 3), f_dc=f(P, 1, 3), f_rest=f(P, M - 1, 3), opacity=f(P, 1), scaling=f(P, 3), rotation=f(P, 4))


def test_header_and_record_layout(tmp_path):
    m = _model()
    path = str(tmp_path / "sub" / "point_cloud.ply")            # save_ply creates the directory (gs_renderer.py:399)
    ply.save_ply(path, *[m[k] for k in ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")])
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().split("\n")
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 37"]
    names = [l.split()[2] for l in lines[3:] if l]
    assert all(l.startswith("property float ") for l in lines[3:] if l)
    assert names == ply.attribute_names(3, 24) and len(names) == 6 + 3 + 24 + 1 + 3 + 4
    rec = np.frombuffer(body, "<f4").reshape(37, len(names))
    assert np.array_equal(rec[:, 0:3], m["xyz"].numpy()) and not rec[:, 3:6].any()
    # channel-major SH: f_rest_0..7 are the 8 red coefficients (features_rest.transpose(1, 2).flatten(1))
    assert np.array_equal(rec[:, 9:17], m["f_rest"].numpy()[:, :, 0]) and np.array_equal(rec[:, 6:9], m["f_dc"].numpy()[:, 0, :])
    assert np.array_equal(rec[:, 33], m["opacity"].numpy()[:, 0]) and np.array_equal(rec[:, -4:], m["rotation"].numpy())


def test_round_trip_all_degrees(tmp_path):
    for deg in range(4):
        m = _model(P=11, deg=deg, seed=deg)
        path = str(tmp_path / ("m%d.ply" % deg))
        ply.save_ply(path, *[m[k] for k in ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")])
        back = ply.load_ply(path, deg, device="cpu")
        for k in m:
            assert back[k].shape == m[k].shape and torch.equal(back[k], m[k]), (deg, k)
    try:
        ply.load_ply(path, 1, device="cpu")
        raise SystemExit("expected the reference's assertion on the f_rest count")
    except AssertionError:
        pass


def test_reads_ascii_and_foreign_property_order(tmp_path):
    path = str(tmp_path / "a.ply")
    names = ["x", "y", "z", "opacity", "f_dc_0", "f_dc_1", "f_dc_2", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3", "extra"]
    rows = np.arange(2 * len(names), dtype=np.float64).reshape(2, -1) / 7
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 2\n" + "".join("property double %s\n" % n for n in names) +
                "element face 0\nproperty list uchar int vertex_indices\nend_header\n")
        for r in rows:
            f.write(" ".join(repr(float(v)) for v in r) + "\n")
    m = ply.load_ply(path, 0, device="cpu")
    assert m["f_rest"].shape == (2, 0, 3) and torch.allclose(m["opacity"][:, 0], torch.tensor(rows[:, 3], dtype=torch.float))
    assert torch.allclose(m["rotation"], torch.tensor(rows[:, 10:14], dtype=torch.float))
