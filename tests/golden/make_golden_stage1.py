"""Golden vectors for SURVEY.md §8 row f2 (optimiser plumbing of the stage-1 loop), made by RUNNING the reference's own
code on the CPU in this container (third-party imports stubbed, see make_golden.py):
  * get_expon_lr_func (gs_renderer.py:25-47) with configs/image.yaml's position-LR settings;
  * GaussianModel.training_setup + two torch.optim.Adam steps + densify_and_prune (gs_renderer.py:356-374, 464-609), with
    torch.normal replaced by recorded standard-normal draws so that the split is reproducible.
Writes tests/golden/stage1_vectors.npz.  Run: python tests/golden/make_golden_stage1.py"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stage1_vectors.npz")
NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def state_of(gm):
    out = {}
    for grp in gm.optimizer.param_groups:
        p = grp["params"][0]
        st = gm.optimizer.state[p]
        out[grp["name"]] = p.detach().numpy().copy()
        out[grp["name"] + "_exp_avg"] = st["exp_avg"].numpy().copy()
        out[grp["name"] + "_exp_avg_sq"] = st["exp_avg_sq"].numpy().copy()
    out["xyz_gradient_accum"] = gm.xyz_gradient_accum.numpy().copy()
    out["denom"] = gm.denom.numpy().copy()
    out["max_radii2D"] = gm.max_radii2D.numpy().copy()
    return out


def main():
    _, gs, _ = make_golden.import_reference()
    out = {}
    f = gs.get_expon_lr_func(lr_init=0.001 * 10, lr_final=0.00002 * 10, lr_delay_mult=0.02, max_steps=500)
    steps = np.arange(0, 521)
    out["lr_steps"], out["lr_values"] = steps, np.array([f(int(s)) for s in steps], np.float64)
    f2 = gs.get_expon_lr_func(lr_init=0.01, lr_final=0.0001, lr_delay_steps=100, lr_delay_mult=0.1, max_steps=300)
    out["lr2_values"] = np.array([f2(int(s)) for s in steps], np.float64)

    rng = np.random.default_rng(7)
    P = 400
    gm = gs.GaussianModel(1)
    scal = np.log(np.exp(rng.normal(-3.4, 0.5, (P, 3))))          # sigma around 0.033: both clone (<= 0.04) and split (> 0.04) happen
    init = dict(xyz=rng.normal(0, 0.3, (P, 3)), f_dc=rng.normal(0, 1, (P, 1, 3)), f_rest=rng.normal(0, 0.1, (P, 3, 3)),
                opacity=rng.normal(-1.0, 2.5, (P, 1)), scaling=scal, rotation=rng.normal(0, 1, (P, 4)))
    t = {k: torch.nn.Parameter(torch.tensor(v, dtype=torch.float32)) for k, v in init.items()}
    gm._xyz, gm._features_dc, gm._features_rest, gm._opacity, gm._scaling, gm._rotation = (t[k] for k in NAMES)
    gm.max_radii2D = torch.zeros(P)
    gm.spatial_lr_scale = 10
    args = types.SimpleNamespace(percent_dense=0.01, position_lr_init=0.001, position_lr_final=0.00002, position_lr_delay_mult=0.02,
                                 position_lr_max_steps=500, feature_lr=0.01, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.005)
    gm.training_setup(args)
    for k, v in init.items():
        out["init_" + k] = np.asarray(v, np.float32)
    grads = []
    for it in (1, 2):
        gm.update_learning_rate(it)
        gstep = {k: (rng.normal(0, 1, init[k].shape) * (10.0 ** rng.integers(-4, 1))).astype(np.float32) for k in NAMES}
        for k in NAMES:
            t[k].grad = torch.tensor(gstep[k])
            out["grad%d_%s" % (it, k)] = gstep[k]
        gm.optimizer.step()
        gm.optimizer.zero_grad()
    for k, v in state_of(gm).items():
        out["after_adam_" + k] = v
    # statistics as the training loop leaves them
    gm.xyz_gradient_accum = torch.tensor(np.abs(rng.normal(0, 0.02, (P, 1))).astype(np.float32) * 3)
    gm.denom = torch.tensor(rng.integers(0, 4, (P, 1)).astype(np.float32))              # zeros -> NaN -> 0 path (:593-594)
    gm.max_radii2D = torch.tensor(rng.integers(0, 3, (P,)).astype(np.float32))
    out["stats_xyz_gradient_accum"], out["stats_denom"], out["stats_max_radii2D"] = (gm.xyz_gradient_accum.numpy().copy(), gm.denom.numpy().copy(),
                                                                                      gm.max_radii2D.numpy().copy())
    noise = rng.normal(0, 1, (4 * P, 3)).astype(np.float32)
    out["split_noise"] = noise
    torch.normal = lambda mean, std: mean + std * torch.tensor(noise[: std.shape[0]])
    gm.densify_and_prune(0.01, min_opacity=0.01, extent=4, max_screen_size=1)
    for k, v in state_of(gm).items():
        out["after_densify_" + k] = v
    print("points: %d -> %d" % (P, gm.get_xyz.shape[0]))
    # reset_opacity (gs_renderer.py:417-420) followed by one more optimiser step: the opacity group restarts from zero
    # moments but keeps its step count (replace_tensor_to_optimizer :464-477 leaves stored_state["step"] alone)
    gm.reset_opacity()
    for k, v in state_of(gm).items():
        out["after_reset_" + k] = v
    gm.update_learning_rate(3)
    n = gm.get_xyz.shape[0]
    shapes = dict(xyz=(n, 3), f_dc=(n, 1, 3), f_rest=(n, 3, 3), opacity=(n, 1), scaling=(n, 3), rotation=(n, 4))
    for grp in gm.optimizer.param_groups:
        g3 = rng.normal(0, 0.1, shapes[grp["name"]]).astype(np.float32)
        out["grad3_" + grp["name"]] = g3
        grp["params"][0].grad = torch.tensor(g3)
    gm.optimizer.step()
    for k, v in state_of(gm).items():
        out["after_reset_adam_" + k] = v
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) >> 10, "KiB")


if __name__ == "__main__":
    main()
