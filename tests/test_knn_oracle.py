"""CPU tests of the distCUDA2 oracle: oracle/knn_oracle.c against scipy's exact kd-tree, closed
forms, and the reference's edge behaviour (/root/reference/simple-knn/simple_knn.cu:118-183)."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

import helpers  # noqa: F401  (puts the repo root on sys.path)
from oracle import knn_oracle

FLT_MAX = np.float32(np.finfo(np.float32).max)


def _kdtree(p):
    d, _ = cKDTree(p.astype(np.float64)).query(p.astype(np.float64), k=4)
    return (d[:, 1:] ** 2).mean(axis=1)


@pytest.mark.parametrize("kind", ["ball", "clusters", "plane", "line"])
def test_oracle_equals_exact_kdtree(kind):
    rng = np.random.default_rng(5)
    P = 4000
    if kind == "ball":
        p = rng.normal(size=(P, 3)); p = p / np.linalg.norm(p, axis=1, keepdims=True) * np.cbrt(rng.random((P, 1))) * 0.5
    elif kind == "clusters":
        p = rng.normal(size=(P, 3)) * 0.01 + rng.integers(0, 5, (P, 1)) * np.array([[1.0, -2.0, 0.5]])
    elif kind == "plane":
        p = np.concatenate([rng.random((P, 2)), np.zeros((P, 1))], axis=1)
    else:
        p = np.concatenate([rng.random((P, 1)), np.zeros((P, 2))], axis=1)
    p = p.astype(np.float32)
    ref = _kdtree(p)
    np.testing.assert_allclose(knn_oracle.dist2_f64(p), ref, rtol=1e-12, atol=0)
    np.testing.assert_allclose(knn_oracle.dist2_f32(p), ref, rtol=2e-6, atol=1e-30)


def test_known_answers_and_small_sets():
    p = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    np.testing.assert_allclose(knn_oracle.dist2_f32(p), [1.0, 5.0 / 3.0, 5.0 / 3.0, 5.0 / 3.0], rtol=1e-7)
    # coincident points count (the reference skips only the query's own index, simple_knn.cu:149,171)
    q = np.array([[0, 0, 0], [0, 0, 0], [2, 0, 0], [0, 2, 0], [0, 0, 2]], np.float32)
    np.testing.assert_allclose(knn_oracle.dist2_f32(q)[:2], [(0 + 4 + 4) / 3.0] * 2, rtol=1e-7)
    # fewer than 4 points: the unfilled slots keep FLT_MAX (simple_knn.cu:139,182)
    with np.errstate(over="ignore"):
        three = knn_oracle.dist2_f32(p[:3])
        assert three[0] == (np.float32(1) + np.float32(1) + FLT_MAX) / np.float32(3)
        assert np.isinf(knn_oracle.dist2_f32(p[:2])).all() and np.isinf(knn_oracle.dist2_f32(p[:1])).all()
    assert knn_oracle.dist2_f32(p[:0]).shape == (0,)


def test_reference_call_site_scales_from_dist2():
    """gs_renderer.py:341-342: dist2 = clamp_min(distCUDA2(points), 1e-7); scales = log(sqrt(dist2)) repeated 3x —
    SURVEY.md §8(d) quotes the resulting mean sigma for the reference's init: 0.0352 at 5k points in a 0.5 ball."""
    from dreamgaussian_b200 import scene
    cloud = scene.make_cloud(5000, 0, seed=0, anisotropic=False)
    dist2 = np.maximum(knn_oracle.dist2_f32(cloud["means3D"]), 1e-7)
    sigma = np.sqrt(dist2)
    assert abs(sigma.mean() - 0.0352) < 0.002
    np.testing.assert_allclose(cloud["scales"][:, 0], sigma, rtol=1e-5)      # scene.make_cloud uses the same definition
