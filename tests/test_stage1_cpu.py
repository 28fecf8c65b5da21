"""CPU tests of the stage-1 optimiser plumbing against outputs of the REFERENCE's own code
(tests/golden/stage1_vectors.npz, made by tests/golden/make_golden_stage1.py from gs_renderer.py:25-47, 356-374, 464-609
and torch.optim.Adam — the optimiser the reference instantiates)."""
import os

import numpy as np
import torch

import helpers  # noqa: F401
from dreamgaussian_b200 import stage1

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stage1_vectors.npz"))
NAMES = stage1.GROUPS


def test_lr_schedule_matches_the_reference_function():
    f = stage1.get_expon_lr_func(0.001 * 10, 0.00002 * 10, lr_delay_mult=0.02, max_steps=500)
    np.testing.assert_allclose([f(int(s)) for s in GOLD["lr_steps"]], GOLD["lr_values"], rtol=1e-13)
    f2 = stage1.get_expon_lr_func(0.01, 0.0001, lr_delay_steps=100, lr_delay_mult=0.1, max_steps=300)
    np.testing.assert_allclose([f2(int(s)) for s in GOLD["lr_steps"]], GOLD["lr2_values"], rtol=1e-13)
    assert stage1.get_expon_lr_func(0.3, 0.3)(17) == 0.3 and stage1.get_expon_lr_func(0.0, 0.0)(5) == 0.0


def _model_from_golden(device="cpu"):
    m = stage1.GaussianModelB200(1)
    m.spatial_lr_scale = 10
    m.fused_adam = device != "cpu"
    m.fused_densify = device != "cpu"
    m._set({k: torch.tensor(GOLD["init_" + k], device=device) for k in NAMES})
    m.training_setup(stage1.OptimConfig())
    return m


def _two_adam_steps(m):
    for it in (1, 2):
        m.update_learning_rate(it)
        for k in NAMES:
            m.p[k].grad = torch.tensor(GOLD["grad%d_%s" % (it, k)], device=m.p[k].device)
        m.optimizer_step()
        m.zero_grad()


def check_against(m, prefix, rtol, atol):
    for k in NAMES:
        for suffix, src in (("", m.p), ("_exp_avg", m.exp_avg), ("_exp_avg_sq", m.exp_avg_sq)):
            want = GOLD[prefix + k + suffix]
            got = src[k].detach().cpu().numpy()
            assert got.shape == want.shape, (k, suffix, got.shape, want.shape)
            np.testing.assert_allclose(got, want, rtol=rtol, atol=atol * max(1e-30, float(np.abs(want).max())), err_msg=k + suffix)


def test_adam_update_matches_torch_optim_adam_as_the_reference_runs_it():
    m = _model_from_golden()
    _two_adam_steps(m)
    check_against(m, "after_adam_", rtol=2e-6, atol=1e-7)


def test_densify_and_prune_matches_the_reference_methods():
    m = _model_from_golden()
    _two_adam_steps(m)
    m.stats.xyz_gradient_accum = torch.tensor(GOLD["stats_xyz_gradient_accum"]).reshape(-1)
    m.stats.denom = torch.tensor(GOLD["stats_denom"]).reshape(-1)
    m.stats.max_radii2D = torch.tensor(GOLD["stats_max_radii2D"]).reshape(-1)
    # the reference drew torch.normal(mean=0, std=stds) = std * noise[:n]; the same draws are handed over
    n_before = m.num_points
    noise = torch.tensor(GOLD["split_noise"])

    class Noise:                                   # sized lazily: the number of split points is known inside
        def to(self, stds):
            return noise[: stds.shape[0]].to(stds)
    m.densify_and_prune(0.01, min_opacity=0.01, extent=4, max_screen_size=1, noise=Noise())
    assert n_before == 400 and m.num_points == GOLD["after_densify_xyz"].shape[0] == 684
    check_against(m, "after_densify_", rtol=3e-6, atol=1e-7)
    # statistics are reset by the densification and then pruned with the points (gs_renderer.py:549-551, 509-512)
    assert not m.stats.xyz_gradient_accum.any() and not m.stats.denom.any() and not m.stats.max_radii2D.any()
    assert m.stats.denom.shape[0] == m.num_points


def test_synthetic_input_and_guidance_stub_contract():
    rgb, mask = stage1.synthetic_rgba(64)
    assert rgb.shape == (1, 3, 64, 64) and mask.shape == (1, 1, 64, 64) and 0.2 < mask.mean() < 0.6
    assert np.allclose(rgb[:, :, 0, 0], 1.0)       # white where the mask is 0 (main.py:100-104)
    g = stage1.GuidanceStub("cpu")
    x = torch.rand((1, 3, 128, 128), requires_grad=True)
    loss = g.train_step(x, [0], [0], [0], step_ratio=0.5)
    loss.backward()
    assert loss.dim() == 0 and x.grad is not None and float(x.grad.abs().sum()) > 0


def test_reset_opacity_matches_the_reference_method_and_keeps_the_step_count():
    """gs_renderer.py:417-420, 464-477 (main.py:285-286): golden = the reference's reset_opacity after its densify_and_prune,
    then one more torch.optim.Adam step."""
    m = _model_from_golden()
    _two_adam_steps(m)
    m.stats.xyz_gradient_accum = torch.tensor(GOLD["stats_xyz_gradient_accum"]).reshape(-1)
    m.stats.denom = torch.tensor(GOLD["stats_denom"]).reshape(-1)
    m.stats.max_radii2D = torch.tensor(GOLD["stats_max_radii2D"]).reshape(-1)
    noise = torch.tensor(GOLD["split_noise"])

    class Noise:
        def to(self, stds):
            return noise[: stds.shape[0]].to(stds)
    m.densify_and_prune(0.01, min_opacity=0.01, extent=4, max_screen_size=1, noise=Noise())
    m.reset_opacity()
    check_against(m, "after_reset_", rtol=3e-6, atol=1e-7)
    assert float(torch.sigmoid(m.p["opacity"]).max()) <= 0.01 + 1e-7 and not m.exp_avg["opacity"].any()
    m.update_learning_rate(3)
    for k in NAMES:
        m.p[k].grad = torch.tensor(GOLD["grad3_" + k])
    m.optimizer_step()
    check_against(m, "after_reset_adam_", rtol=5e-6, atol=1e-7)
