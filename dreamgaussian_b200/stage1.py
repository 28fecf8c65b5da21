"""the stage-1 optimisation loop around the rasterizer (BASELINE.json configs[3]) — GaussianModel's
optimiser plumbing restated B200-first.

Reference: /root/reference/gs_renderer.py:25-47 (LR schedule), :331-382 (create_from_pcd, training_setup,
update_learning_rate), :464-622 (optimizer-state surgery, densify_and_clone / _split / _prune, prune),
main.py:182-287 (train_step).  What changes:
  * the six parameter tensors are stepped by ONE fused Adam launch (dgr_adam_step) instead of torch.optim.Adam's
    six groups x ~6 element-wise kernels; Adam state lives in plain tensors, so the reference's optimizer-state surgery
    (replace / prune / cat on `optimizer.state`) becomes ordinary tensor indexing;
  * rendering goes through FusedGaussianRasterizer (row f1): activations, SH dc/rest split and the three densification
    statistics happen inside the per-Gaussian kernels;
  * initial scales come from this library's distCUDA2 (row f3).
The densification exists twice: as device-agnostic tensor logic (tests/test_stage1_cpu.py runs it on the CPU against outputs
of the reference's own methods) and as stream-compaction kernels (dgr_densify_plan / dgr_densify_apply, csrc/dgr_densify.cuh)
that the CUDA loop uses and tests/test_stage1_gpu.py pins against the same reference outputs."""
import ctypes
import math
from typing import NamedTuple, Optional

import numpy as np
import torch

from . import _lib
from .fused import DensifyStats, FusedGaussianRasterizer
from .rasterizer import GaussianRasterizationSettings

SH_C0 = 0.28209479177387814


def get_expon_lr_func(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    """gs_renderer.py:25-47: log-linear decay from lr_init to lr_final with an optional sine warm-up."""
    def helper(step):
        if lr_init == lr_final:
            return lr_init
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        if lr_delay_steps > 0:
            delay_rate = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0.0), 1.0))
        else:
            delay_rate = 1.0
        t = min(max(step / max_steps, 0.0), 1.0)
        return delay_rate * math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)
    return helper


class OptimConfig(NamedTuple):
    """configs/image.yaml:60-80"""
    position_lr_init: float = 0.001
    position_lr_final: float = 0.00002
    position_lr_delay_mult: float = 0.02
    position_lr_max_steps: int = 500
    feature_lr: float = 0.01
    opacity_lr: float = 0.05
    scaling_lr: float = 0.005
    rotation_lr: float = 0.005
    percent_dense: float = 0.01


GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def build_rotation(r):
    """gs_renderer.py:85-107 (normalises the quaternion)."""
    q = r / torch.sqrt((r * r).sum(dim=1, keepdim=True))
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=1).reshape(-1, 3, 3)


class GaussianModelB200:
    """The reference GaussianModel's training state: raw parameters (same attribute names), Adam moments, statistics."""

    def __init__(self, sh_degree: int):
        self.max_sh_degree = sh_degree
        self.active_sh_degree = sh_degree          # main.py:147: no progressive SH
        self.p = {}                                # name -> leaf tensor (requires_grad)
        self.exp_avg, self.exp_avg_sq = {}, {}
        self.lr = {}
        self.steps = {k: 0 for k in GROUPS}          # torch.optim.Adam keeps state['step'] per tensor
        self.stats: Optional[DensifyStats] = None
        self.percent_dense = 0.01
        self.spatial_lr_scale = 1.0
        self.betas, self.eps = (0.9, 0.999), 1e-15
        self.fused_adam = True

    # reference attribute names (gs_renderer.py:140-160) so that fused.render_gaussian_model / fields work on this object
    _xyz = property(lambda s: s.p["xyz"]); _features_dc = property(lambda s: s.p["f_dc"]); _features_rest = property(lambda s: s.p["f_rest"])
    _opacity = property(lambda s: s.p["opacity"]); _scaling = property(lambda s: s.p["scaling"]); _rotation = property(lambda s: s.p["rotation"])

    @property
    def num_points(self):
        return self.p["xyz"].shape[0]

    def create_from_points(self, points, colors, spatial_lr_scale=1.0, device="cuda"):
        """gs_renderer.py:331-354 (colors in [0,1], SH dc = RGB2SH; scales from the 3-NN mean squared distance)."""
        from .knn import distCUDA2
        self.spatial_lr_scale = spatial_lr_scale
        xyz = torch.tensor(np.asarray(points), dtype=torch.float32, device=device)
        rgb = torch.tensor(np.asarray(colors), dtype=torch.float32, device=device)
        M = (self.max_sh_degree + 1) ** 2
        dist2 = torch.clamp_min(distCUDA2(xyz), 0.0000001)
        self._set(dict(xyz=xyz, f_dc=((rgb - 0.5) / SH_C0).reshape(-1, 1, 3).contiguous(),
                       f_rest=torch.zeros((xyz.shape[0], M - 1, 3), device=device),
                       opacity=torch.full((xyz.shape[0], 1), math.log(0.1 / 0.9), device=device),
                       scaling=torch.log(torch.sqrt(dist2))[..., None].repeat(1, 3),
                       rotation=torch.tensor([[1.0, 0, 0, 0]], device=device).repeat(xyz.shape[0], 1)))

    def _set(self, tensors):
        self.p = {k: tensors[k].detach().float().contiguous().requires_grad_(True) for k in GROUPS}

    def training_setup(self, cfg: OptimConfig = OptimConfig()):
        """gs_renderer.py:356-374"""
        self.percent_dense = cfg.percent_dense
        self._reset_stats()
        s = self.spatial_lr_scale
        self.lr = dict(xyz=cfg.position_lr_init * s, f_dc=cfg.feature_lr, f_rest=cfg.feature_lr / 20.0, opacity=cfg.opacity_lr,
                       scaling=cfg.scaling_lr, rotation=cfg.rotation_lr)
        self.exp_avg = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.exp_avg_sq = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.steps = {k: 0 for k in GROUPS}
        self.xyz_scheduler = get_expon_lr_func(cfg.position_lr_init * s, cfg.position_lr_final * s, lr_delay_mult=cfg.position_lr_delay_mult,
                                               max_steps=cfg.position_lr_max_steps)

    def _reset_stats(self):
        self.stats = DensifyStats(self.num_points, self.p["xyz"].device)

    def update_learning_rate(self, iteration):
        self.lr["xyz"] = self.xyz_scheduler(iteration)
        return self.lr["xyz"]

    def zero_grad(self):
        for v in self.p.values():
            v.grad = None

    def optimizer_step(self):
        """torch.optim.Adam.step() of the reference (main.py:275) as one launch; a group whose .grad is None is skipped,
        as torch does."""
        names = [k for k in GROUPS if self.p[k].grad is not None]
        for k in names:
            self.steps[k] += 1
        names = [k for k in names if self.p[k].numel() > 0]
        if not names:
            return
        if not self.fused_adam or not self.p["xyz"].is_cuda:
            return self._adam_torch(names)
        lib = _lib.load()
        arr = (_lib.DgrAdamGroup * len(names))()
        keep = []
        for i, k in enumerate(names):
            g = self.p[k].grad.contiguous(); keep.append(g)
            arr[i] = _lib.DgrAdamGroup(self.p[k].data_ptr(), g.data_ptr(), self.exp_avg[k].data_ptr(), self.exp_avg_sq[k].data_ptr(),
                                       self.p[k].numel(), self.lr[k], self.steps[k])
        dev = self.p["xyz"].device
        with torch.cuda.device(dev):
            st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib.dgr_adam_step(arr, len(names), self.betas[0], self.betas[1], self.eps, st))

    def _adam_torch(self, names):
        """The same update with torch ops (device-agnostic; used by the CPU tests of the host logic)."""
        b1, b2 = self.betas
        with torch.no_grad():
            for k in names:
                bc1, bc2 = 1 - b1 ** self.steps[k], 1 - b2 ** self.steps[k]
                g = self.p[k].grad
                self.exp_avg[k].lerp_(g, 1 - b1)
                self.exp_avg_sq[k].mul_(b2).addcmul_(g, g, value=1 - b2)
                denom = (self.exp_avg_sq[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
                self.p[k].addcdiv_(self.exp_avg[k], denom, value=-self.lr[k] / bc1)

    # ---- densification (gs_renderer.py:464-622): optimizer-state surgery == tensor indexing here
    def _select(self, mask):
        with torch.no_grad():
            self.p = {k: v[mask].detach().contiguous().requires_grad_(True) for k, v in self.p.items()}
            self.exp_avg = {k: v[mask].contiguous() for k, v in self.exp_avg.items()}
            self.exp_avg_sq = {k: v[mask].contiguous() for k, v in self.exp_avg_sq.items()}
            st = self.stats
            st.xyz_gradient_accum, st.denom, st.max_radii2D = st.xyz_gradient_accum[mask], st.denom[mask], st.max_radii2D[mask]

    def _append(self, new):
        """densification_postfix + cat_tensors_to_optimizer (:515-552): new points get zero Adam moments, stats reset."""
        with torch.no_grad():
            self.p = {k: torch.cat((self.p[k].detach(), new[k]), dim=0).contiguous().requires_grad_(True) for k in GROUPS}
            self.exp_avg = {k: torch.cat((v, torch.zeros_like(new[k])), dim=0) for k, v in self.exp_avg.items()}
            self.exp_avg_sq = {k: torch.cat((v, torch.zeros_like(new[k])), dim=0) for k, v in self.exp_avg_sq.items()}
        self._reset_stats()

    def prune_points(self, mask):
        self._select(~mask)

    def densify_and_clone(self, grads, grad_threshold, scene_extent):
        with torch.no_grad():
            sel = (grads >= grad_threshold) & (torch.exp(self.p["scaling"]).max(dim=1).values <= self.percent_dense * scene_extent)
            self._append({k: self.p[k].detach()[sel] for k in GROUPS})

    def densify_and_split(self, grads, grad_threshold, scene_extent, N=2, noise=None):
        """:555-580.  noise: optional standard-normal [N * n_selected, 3] (tests pass the reference's draws)."""
        with torch.no_grad():
            n_init = self.num_points
            padded = torch.zeros((n_init,), device=grads.device)
            padded[:grads.shape[0]] = grads
            scaling = torch.exp(self.p["scaling"].detach())
            sel = (padded >= grad_threshold) & (scaling.max(dim=1).values > self.percent_dense * scene_extent)
            stds = scaling[sel].repeat(N, 1)
            if noise is None:
                noise = torch.randn(stds.shape, device=stds.device)
            samples = noise.to(stds) * stds
            rots = build_rotation(self.p["rotation"].detach()[sel]).repeat(N, 1, 1)
            new = dict(xyz=torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + self.p["xyz"].detach()[sel].repeat(N, 1),
                       scaling=torch.log(scaling[sel].repeat(N, 1) / (0.8 * N)),
                       rotation=self.p["rotation"].detach()[sel].repeat(N, 1),
                       f_dc=self.p["f_dc"].detach()[sel].repeat(N, 1, 1), f_rest=self.p["f_rest"].detach()[sel].repeat(N, 1, 1),
                       opacity=self.p["opacity"].detach()[sel].repeat(N, 1))
            n_new = new["xyz"].shape[0]
            self._append(new)
            self.prune_points(torch.cat((sel, torch.zeros((n_new,), dtype=torch.bool, device=sel.device))))

    def _prune_mask(self, min_opacity, extent, max_screen_size):
        m = (torch.sigmoid(self.p["opacity"].detach()) < min_opacity).squeeze(1)
        if max_screen_size:
            m = m | (self.stats.max_radii2D > max_screen_size) | (torch.exp(self.p["scaling"].detach()).max(dim=1).values > 0.1 * extent)
        return m

    def _densify_fused(self, max_grad, min_opacity, extent, max_screen_size, noise=None):
        """densify_and_prune as two launches of compaction kernels (csrc/dgr_densify.cuh: classify + block scan, then ONE pass
        that writes the six parameter tensors and both Adam moments in the reference's final order) with one host read-back
        of the new point count between them — instead of ~30 boolean-mask indexing / torch.cat calls."""
        lib = _lib.load()
        P, dev = self.num_points, self.p["xyz"].device
        with torch.no_grad(), torch.cuda.device(dev):
            st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            scratch = torch.empty((lib.dgr_densify_scratch_bytes(P),), dtype=torch.uint8, device=dev)
            if getattr(self, "_counts_host", None) is None:
                self._counts_host = torch.zeros((4,), dtype=torch.int32).pin_memory()
            acc, den = self.stats.xyz_gradient_accum.contiguous(), self.stats.denom.contiguous()
            _lib.check(lib.dgr_densify_plan(P, acc.data_ptr(), den.data_ptr(), self.p["opacity"].data_ptr(), self.p["scaling"].data_ptr(),
                                            float(max_grad), float(self.percent_dense * extent), float(min_opacity), float(0.1 * extent),
                                            1 if max_screen_size else 0, scratch.data_ptr(), self._counts_host.data_ptr(), st))
            torch.cuda.current_stream(dev).synchronize()
            n_keep, n_clone, n_sel, n_child = (int(x) for x in self._counts_host)
            n_out = n_keep + n_clone + 2 * n_child
            if noise is None:
                noise = torch.randn((max(2 * n_sel, 1), 3), device=dev)
            else:                                   # tests hand over the reference's draws: row = child * n_selected + rank
                noise = noise.to(torch.empty((2 * n_sel, 3), device=dev)) if hasattr(noise, "to") else noise
            noise = noise.contiguous().float()
            t = _lib.DgrDensifyTensors()
            new_p, new_m, new_v, keep = {}, {}, {}, [noise, scratch]
            for i, k in enumerate(GROUPS):
                src = self.p[k].detach()
                w = int(src[0].numel()) if src.shape[0] else int(np.prod(src.shape[1:]))
                new_p[k] = torch.empty((n_out,) + tuple(src.shape[1:]), dtype=torch.float32, device=dev)
                new_m[k], new_v[k] = torch.empty_like(new_p[k]), torch.empty_like(new_p[k])
                t.inp[i], t.exp_avg_in[i], t.exp_avg_sq_in[i] = src.data_ptr(), self.exp_avg[k].data_ptr(), self.exp_avg_sq[k].data_ptr()
                t.out[i], t.exp_avg_out[i], t.exp_avg_sq_out[i] = new_p[k].data_ptr(), new_m[k].data_ptr(), new_v[k].data_ptr()
                t.width[i] = w
            _lib.check(lib.dgr_densify_apply(P, ctypes.byref(t), noise.data_ptr(), scratch.data_ptr(), st))
            self.p = {k: v.requires_grad_(True) for k, v in new_p.items()}
            self.exp_avg, self.exp_avg_sq = new_m, new_v
        self._reset_stats()

    def densify_and_prune(self, max_grad, min_opacity, extent, max_screen_size, noise=None):
        if getattr(self, "fused_densify", False) and self.p["xyz"].is_cuda and self.num_points > 0:
            return self._densify_fused(max_grad, min_opacity, extent, max_screen_size, noise=noise)
        with torch.no_grad():
            grads = self.stats.xyz_gradient_accum / self.stats.denom
            grads[grads.isnan()] = 0.0
        self.densify_and_clone(grads, max_grad, extent)
        self.densify_and_split(grads, max_grad, extent, noise=noise)
        self.prune_points(self._prune_mask(min_opacity, extent, max_screen_size))

    def prune(self, min_opacity, extent, max_screen_size):
        self.prune_points(self._prune_mask(min_opacity, extent, max_screen_size))

    def reset_opacity(self):
        """gs_renderer.py:417-420 + replace_tensor_to_optimizer :464-477: opacity = inverse_sigmoid(min(sigmoid(o), 0.01)) and
        that group's Adam moments zeroed (torch keeps the group's step count: `stored_state["step"]` is left alone)."""
        with torch.no_grad():
            o = torch.sigmoid(self.p["opacity"].detach())
            o = torch.min(o, torch.full_like(o, 0.01))
            self.p["opacity"] = torch.log(o / (1 - o)).contiguous().requires_grad_(True)
            self.exp_avg["opacity"] = torch.zeros_like(self.p["opacity"])
            self.exp_avg_sq["opacity"] = torch.zeros_like(self.p["opacity"])


# ---------------------------------------------------------------------------------------------------------------------
class Stage1Config(NamedTuple):
    """configs/image.yaml + main.py defaults used by train_step."""
    iters: int = 500
    ref_size: int = 256
    num_pts: int = 5000
    sh_degree: int = 0
    radius: float = 2.0
    fovy: float = 49.1
    elevation: float = 0.0
    min_ver: int = -30
    max_ver: int = 30
    invert_bg_prob: float = 0.5
    lambda_zero123: float = 1.0
    density_start_iter: int = 100
    density_end_iter: int = 3000
    densification_interval: int = 100
    opacity_reset_interval: int = 700
    densify_grad_threshold: float = 0.01
    seed: int = 0


class GuidanceStub:
    """Deterministic stand-in for guidance.zero123_utils.Zero123.train_step: a differentiable
    scalar 0.5 * sum((downsample(pred, 32x32) - fixed_target)^2), as the real one returns 0.5 * ||latents - target||^2."""

    def __init__(self, device, seed=0):
        g = torch.Generator(device="cpu").manual_seed(seed + 99)
        self.target = torch.rand((1, 3, 32, 32), generator=g).to(device)

    def train_step(self, images, vers, hors, radii, step_ratio=None, default_elevation=0):
        pred = torch.nn.functional.interpolate(images, (32, 32), mode="bilinear", align_corners=False)
        return 0.5 * ((pred - self.target) ** 2).sum()


def synthetic_rgba(size):
    """Synthetic RGBA input (BASELINE.json configs[3]): a shaded disc on a transparent background, [1,3,S,S] and [1,1,S,S]."""
    y, x = np.mgrid[0:size, 0:size].astype(np.float32)
    c = (size - 1) / 2.0
    r = np.sqrt((x - c) ** 2 + (y - c) ** 2) / (0.35 * size)
    mask = (r < 1.0).astype(np.float32)
    shade = np.clip(1.0 - 0.6 * r, 0, 1)
    rgb = np.stack([0.9 * shade, 0.4 + 0.3 * shade, 0.2 + 0.5 * (x / size)], 0) * mask + (1 - mask)      # white background (main.py:100-104)
    return rgb[None].astype(np.float32), mask[None, None].astype(np.float32)


class Stage1Trainer:
    """main.py:182-287 train_step, guidance stubbed.  `fused=True`: FusedGaussianRasterizer + fused Adam (rows f1 + f2);
    `fused=False`: the reference's formulation through this library's plain op — torch activations + torch.cat, the three
    torch statistic updates, torch.optim.Adam-style per-tensor updates — the baseline tools/stage1_bench.py times."""

    def __init__(self, cfg: Stage1Config = Stage1Config(), optim: OptimConfig = OptimConfig(), device="cuda", fused=True):
        from . import scene
        self.cfg, self.device, self.fused = cfg, torch.device(device), fused
        self.rng = np.random.default_rng(cfg.seed)
        self.scene = scene
        cloud = scene.make_cloud(cfg.num_pts, cfg.sh_degree, seed=cfg.seed, anisotropic=False, sigma=1.0)      # positions only
        colors = self.rng.random((cfg.num_pts, 3)) / 255.0 * SH_C0 + 0.5                                         # SH2RGB(rand/255), gs_renderer.py:703-706
        self.gaussians = GaussianModelB200(cfg.sh_degree)
        self.gaussians.fused_adam = fused
        self.gaussians.fused_densify = fused
        self.gaussians.create_from_points(cloud["means3D"], colors, spatial_lr_scale=10.0, device=device)      # gs_renderer.py:709
        self.gaussians.training_setup(optim)
        rgb, mask = synthetic_rgba(cfg.ref_size)
        self.input_img, self.input_mask = torch.tensor(rgb, device=device), torch.tensor(mask, device=device)
        self.guidance = GuidanceStub(self.device, cfg.seed)
        self.step = 0
        self.fixed_cam = self._settings(cfg.elevation, 0.0, cfg.radius, cfg.ref_size, (1.0, 1.0, 1.0))
        self.losses = []

    def _settings(self, elev, azim, radius, res, bg):
        cam = self.scene.orbit_camera(elev, azim, radius, res, res, fovy_deg=self.cfg.fovy)
        # one host-to-device copy for the four small camera tensors (the loop is host-bound: every copy is a driver call)
        buf = np.concatenate([np.append(np.asarray(bg, np.float32).ravel(), 0.0).astype(np.float32), cam.world_view_transform.ravel(), cam.full_proj_transform.ravel(),
                              cam.camera_center.ravel(), np.zeros(1, np.float32)]).astype(np.float32)
        t = torch.from_numpy(buf).to(self.device)
        return GaussianRasterizationSettings(image_height=res, image_width=res, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=t[0:3],
                                             scale_modifier=1.0, viewmatrix=t[4:20].view(4, 4), projmatrix=t[20:36].view(4, 4),
                                             sh_degree=self.gaussians.active_sh_degree, campos=t[36:39], prefiltered=False, debug=False)

    def render(self, rs, track_stats):
        g = self.gaussians
        means2D = torch.zeros_like(g._xyz, requires_grad=True)
        if self.fused:
            color, radii, depth, alpha = FusedGaussianRasterizer(rs)(g._xyz, g._features_dc, g._features_rest, g._opacity, g._scaling,
                                                                     g._rotation, means2D=means2D, stats=g.stats if track_stats else None)
        else:
            from .rasterizer import GaussianRasterizer
            color, radii, depth, alpha = GaussianRasterizer(rs)(
                means3D=g._xyz, means2D=means2D, shs=torch.cat((g._features_dc, g._features_rest), dim=1),
                opacities=torch.sigmoid(g._opacity), scales=torch.exp(g._scaling), rotations=torch.nn.functional.normalize(g._rotation))
        return {"image": color.clamp(0, 1), "alpha": alpha, "viewspace_points": means2D, "visibility_filter": radii > 0, "radii": radii}

    def train_step(self):
        cfg, g = self.cfg, self.gaussians
        self.step += 1
        step_ratio = min(1, self.step / cfg.iters)
        g.update_learning_rate(self.step)
        g.zero_grad()
        densify_window = cfg.density_start_iter <= self.step <= cfg.density_end_iter
        # known view (main.py:196-208)
        out = self.render(self.fixed_cam, track_stats=False)
        loss = 10000 * step_ratio * torch.nn.functional.mse_loss(out["image"].unsqueeze(0), self.input_img)
        loss = loss + 1000 * step_ratio * torch.nn.functional.mse_loss(out["alpha"].unsqueeze(0), self.input_mask)
        # novel view (main.py:210-240); the statistics come from this LAST render of the step (main.py:278-281)
        res = 128 if step_ratio < 0.3 else (256 if step_ratio < 0.6 else 512)
        min_ver = max(min(cfg.min_ver, cfg.min_ver - cfg.elevation), -80 - cfg.elevation)
        max_ver = min(max(cfg.max_ver, cfg.max_ver - cfg.elevation), 80 - cfg.elevation)
        ver, hor = int(self.rng.integers(min_ver, max_ver)), int(self.rng.integers(-180, 180))
        bg = (1.0, 1.0, 1.0) if self.rng.random() > cfg.invert_bg_prob else (0.0, 0.0, 0.0)
        out = self.render(self._settings(cfg.elevation + ver, hor, cfg.radius, res, bg), track_stats=densify_window)
        loss = loss + cfg.lambda_zero123 * self.guidance.train_step(out["image"].unsqueeze(0), [ver], [hor], [0], step_ratio=step_ratio,
                                                                    default_elevation=cfg.elevation)
        loss.backward()
        g.optimizer_step()
        if densify_window:
            if not self.fused:          # the reference's three torch updates (main.py:279-281, gs_renderer.py:625-627)
                with torch.no_grad():
                    vis, radii, st = out["visibility_filter"], out["radii"], g.stats
                    st.max_radii2D[vis] = torch.max(st.max_radii2D[vis], radii[vis].float())
                    st.xyz_gradient_accum[vis] += torch.norm(out["viewspace_points"].grad[vis, :2], dim=-1)
                    st.denom[vis] += 1
            if self.step % cfg.densification_interval == 0:
                g.densify_and_prune(cfg.densify_grad_threshold, min_opacity=0.01, extent=4, max_screen_size=1)
            if self.step % cfg.opacity_reset_interval == 0:          # main.py:285-286 (700 > 500 iterations in image.yaml)
                g.reset_opacity()
        self.losses.append(loss.detach())
        return loss

    def train(self, iters=None):
        for _ in range(iters or self.cfg.iters):
            self.train_step()
        return torch.stack(self.losses).cpu().numpy()
