"""GPU parity of extract_fields: dgr_extract_fields (through dreamgaussian_b200.fields -> C ABI)
against the reference's own outputs (golden fixtures) and against the CPU oracle on larger seeded inputs.
Tolerance: FIELD_ATOL * max(1, max occ) — see tests/test_fields_oracle.py."""
import os

import numpy as np
import pytest
import torch

import helpers  # noqa: F401
from dreamgaussian_b200 import fields, scene
from oracle import fields_oracle

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "extract_fields_vectors.npz"))
FIELD_ATOL = 5e-5


def _run(args, res, nb, relax):
    t = [torch.tensor(a, device="cuda") for a in args]
    occ, center, scale = fields.extract_fields(*t, resolution=res, num_blocks=nb, relax_ratio=relax)
    return occ.cpu().numpy(), center.cpu().numpy(), float(scale)


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_reproduces_the_reference_outputs(case):
    res, nb, relax = GOLD[case + "_params"]
    args = [GOLD["%s_%s" % (case, k)] for k in ("xyz", "opacity", "scaling", "rotation")]
    occ, center, scale = _run(args, int(res), int(nb), float(relax))
    ref = GOLD[case + "_occ"]
    assert np.abs(occ - ref).max() <= FIELD_ATOL * max(1.0, float(ref.max()))
    assert (occ == 0).sum() >= 0.98 * (ref == 0).sum()           # the per-block truncation leaves the same voxels empty
    assert np.array_equal(center, GOLD[case + "_center"]) and scale == np.float32(GOLD[case + "_scale"])


@pytest.mark.parametrize("P,res,nb", [(20000, 64, 16), (100000, 128, 16), (5000, 96, 8), (3000, 64, 4)])
def test_matches_the_oracle_on_a_model_like_cloud(P, res, nb):
    raw = scene.to_raw_parameters(scene.make_cloud(P, 0, seed=4, sigma=None if P <= 20000 else 0.0128))
    args = [raw["xyz"], raw["opacity"], raw["scaling"], raw["rotation"]]
    want, wc, ws = fields_oracle.extract_fields(*args, res, nb, 1.5)
    occ, center, scale = _run(args, res, nb, 1.5)
    assert np.abs(occ - want).max() <= FIELD_ATOL * max(1.0, float(want.max()))
    assert np.array_equal(center, wc) and scale == np.float32(ws)


def test_model_style_call_and_validation():
    class G:
        pass
    g = G()
    raw = scene.to_raw_parameters(scene.make_cloud(2000, 0, seed=1, sigma=0.03))
    g._xyz, g._opacity, g._scaling, g._rotation = (torch.tensor(raw[k], device="cuda") for k in ("xyz", "opacity", "scaling", "rotation"))
    occ = fields.extract_fields_of_model(g, resolution=32)
    assert occ.shape == (32, 32, 32) and isinstance(g.scale, float) and g.center.shape == (3,)
    with pytest.raises(ValueError):
        fields.extract_fields(g._xyz, g._opacity, g._scaling, g._rotation, resolution=30, num_blocks=16)
    with pytest.raises(RuntimeError, match="no CPU path"):
        fields.extract_fields(g._xyz.cpu(), g._opacity, g._scaling, g._rotation)
