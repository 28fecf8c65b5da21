#!/usr/bin/env python
"""bench.py — splats/sec of the differentiable Gaussian-splat rasterize (forward + backward), BASELINE.json's metric.

A "step" = one forward + backward pass of the rasterizer over ONE synthetic view per GPU (views are the data-parallel
axis); with N GPUs every rank renders its own view of the same replicated cloud, accumulates the
per-Gaussian gradients into one flat buffer and the ranks all-reduce it over NCCL (weak scaling).
Workload at N=1: BASELINE.json configs[1] — 100k Gaussians, 800x800, SH degree 3 (synthetic DreamGaussian-like cloud,
dreamgaussian_b200/scene.py).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our sm_100a path (the product)
  python bench.py --impl reference ...                           # the CPU oracle port, timed on the host cores

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=100000)
    ap.add_argument("--res", type=int, default=800)
    ap.add_argument("--sh-degree", type=int, default=3)
    ap.add_argument("--opacity", default="trained", choices=["trained", "init"])
    ap.add_argument("--views", type=int, default=8, help="size of the fixed camera set that steps cycle through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-rows", action="store_true", help="skip the short measurements of the SURVEY §8f rows (f1-f4)")
    ap.add_argument("--seed", type=int, default=0)
    return ap.parse_args()


def workload_name(a):
    return "%dk Gaussians, %dx%d, SH degree %d, forward+backward, opacity=%s, anisotropic" % (
        a.points // 1000, a.res, a.res, a.sh_degree, a.opacity)


def algorithmic_bytes(P, M, H, W, n_inst):
    """B_alg = P (148 + 36 M) + H W 52 + N_inst 44 bytes per view, forward + backward."""
    return P * (148 + 36 * M) + H * W * 52 + n_inst * 44


def kernel_algorithmic_bytes(name, P, M, H, W, n_inst):
    """Per-kernel algorithmic bytes (what the kernel must read + write once), DESIGN.md §Kernels."""
    px = H * W
    table = {
        "preprocess_fwd": P * (44 + 12 * M) + P * (4 + 48 + 4) + n_inst * 4, # inputs; radii + record + touched; histogram
        "tile_scan": (H // 16 + 1) * (W // 16 + 1) * 16,                    # counts in; ranges + cursor out
        "emit_instances": P * 52 + n_inst * 12,                             # record + touched; cursor atomics + key
        "tile_sort_gather": n_inst * (8 + 4 + 48 + 48),                     # key in; id out; record gather + sorted record
        "tile_sort_gather_big": n_inst * (8 + 4 + 48 + 48),                 # the big-tile walker does the same per instance (upper bound: all instances)
        "render_fwd": n_inst * 48 + px * 28,                                 # sorted records; rgb+depth+alpha+n_contrib+T
        "render_bwd": n_inst * 52 + px * 28 + P * 48,                        # records+ids; grads+n_contrib+T; moments
        "preprocess_bwd": P * (44 + 12 * M) + P * 52 + P * (56 + 12 * M),   # inputs; moments+radii; grads
    }
    return table.get(name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def wait_first_sample(self, timeout=3.0):
        """nvidia-smi needs ~100 ms to start: block until it has produced a row so a short run is still covered."""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        return len(self.rows)

    def stop(self, window=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows, where = self.rows, "whole run (timed region shorter than the 20 ms sampling period)"
        if window and window[1] - window[0] >= 3:
            rows, where = self.rows[window[0]:window[1]], "timed region"
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": where}


def build_scene(a):
    from dreamgaussian_b200 import scene
    cloud = scene.make_cloud(a.points, a.sh_degree, seed=a.seed, opacity=a.opacity, anisotropic=True)
    cams = scene.bench_views(a.views, a.res, a.res)
    rng = np.random.default_rng(a.seed + 17)
    ups = [(rng.normal(size=(3, a.res, a.res)).astype(np.float32), None, rng.normal(size=(1, a.res, a.res)).astype(np.float32))
           for _ in range(min(a.views, 4))]
    return cloud, cams, ups


def cam_settings(cam, a, bg):
    return dict(image_height=a.res, image_width=a.res, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0,
                viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, sh_degree=a.sh_degree,
                campos=cam.camera_center)


def host_threads():
    """Threads for the CPU legs: the physical cores this process may run on.  (One OpenMP thread per LOGICAL CPU made the
    oracle 10x slower on the 2 x 32-core / 128-thread boxes; a cgroup CPU quota, if any, caps it further.)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        import psutil
        phys = psutil.cpu_count(logical=False)
        if phys:
            n = min(n, phys)
    except Exception:
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        pass
    return max(1, n)


def measure_next_rows(dev):
    """Short, guarded measurements of the rows next to the hot path — reported beside the
    headline metric, never part of it.  Each entry: this library's time and what it is compared with."""
    import numpy as np
    import torch
    from dreamgaussian_b200 import fields, scene, stage1
    from dreamgaussian_b200.fused import DensifyStats, FusedGaussianRasterizer
    from dreamgaussian_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    from simple_knn._C import distCUDA2
    rows = {}

    def ev_median(fn, n):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        evs = []
        for i in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(i); e1.record(); evs.append((e0, e1))
        torch.cuda.synchronize()
        return float(np.median([x.elapsed_time(y) for x, y in evs]))

    try:        # f1: raw-parameter step (fwd + bwd + densification statistics) at the bench shape
        P, res = 100000, 800
        cloud = scene.make_cloud(P, 3, seed=0, opacity="trained", anisotropic=True)
        raw = {k: torch.tensor(v, device=dev).requires_grad_(True) for k, v in scene.to_raw_parameters(cloud).items()}
        t = lambda x: torch.tensor(np.asarray(x, np.float32), device=dev)
        cams = scene.bench_views(4, res, res)
        rs = [GaussianRasterizationSettings(image_height=res, image_width=res, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=t(np.ones(3)),
              scale_modifier=1.0, viewmatrix=t(c.world_view_transform), projmatrix=t(c.full_proj_transform), sh_degree=3,
              campos=t(c.camera_center), prefiltered=False, debug=False) for c in cams]
        rng = np.random.default_rng(17)
        gC, gA = t(rng.normal(size=(3, res, res))), t(rng.normal(size=(1, res, res)))
        stats = DensifyStats(P, dev)
        acc, den, mr = torch.zeros((P, 1), device=dev), torch.zeros((P, 1), device=dev), torch.zeros((P,), device=dev)

        def zero():
            for v in raw.values():
                v.grad = None

        def ref_form(i):
            zero()
            m2d = torch.zeros_like(raw["xyz"], requires_grad=True)
            c, r, d, al = GaussianRasterizer(rs[i % 4])(means3D=raw["xyz"], means2D=m2d, shs=torch.cat((raw["features_dc"], raw["features_rest"]), dim=1),
                                                       opacities=torch.sigmoid(raw["opacity"]), scales=torch.exp(raw["scaling"]),
                                                       rotations=torch.nn.functional.normalize(raw["rotation"]))
            torch.autograd.backward([c, al], [gC, gA])
            with torch.no_grad():
                vis = r > 0
                mr[vis] = torch.max(mr[vis], r[vis].float()); acc[vis] += torch.norm(m2d.grad[vis, :2], dim=-1, keepdim=True); den[vis] += 1

        def fused(i):
            zero()
            c, r, d, al = FusedGaussianRasterizer(rs[i % 4])(raw["xyz"], raw["features_dc"], raw["features_rest"], raw["opacity"], raw["scaling"],
                                                            raw["rotation"], stats=stats)
            torch.autograd.backward([c, al], [gC, gA])
        rows["f1_raw_parameter_step"] = {"ours_ms": ev_median(fused, 40), "torch_activations_plus_plain_op_ms": ev_median(ref_form, 40),
                                         "shape": "100k / 800x800 / deg 3, fwd+bwd+densify stats through autograd"}
        del raw, stats
    except Exception as e:  # noqa: BLE001
        rows["f1_raw_parameter_step"] = {"error": repr(e)[:200]}
    try:        # f2: BASELINE.json configs[3]
        out = {}
        for name, fz in (("ours", True), ("reference_formulation", False)):
            stage1.Stage1Trainer(stage1.Stage1Config(), fused=fz).train(10)
            tr = stage1.Stage1Trainer(stage1.Stage1Config(), fused=fz)
            torch.cuda.synchronize(); t0 = time.perf_counter(); tr.train(500); torch.cuda.synchronize()
            out[name + "_s_per_500_iters"] = time.perf_counter() - t0
            out[name + "_final_points"] = tr.gaussians.num_points
        out["shape"] = "configs/image.yaml stage-1 loop, synthetic RGBA, guidance stubbed, 500 iterations"
        rows["f2_stage1_loop"] = out
    except Exception as e:  # noqa: BLE001
        rows["f2_stage1_loop"] = {"error": repr(e)[:200]}
    try:        # f3: distCUDA2 against the reference's own CUDA code (oracle/_ref), 100k points of the reference's init ball
        from oracle import knn_oracle
        pts = torch.tensor(scene.make_cloud(100000, 0, seed=1, anisotropic=False, sigma=1.0)["means3D"], device=dev)
        r = {"ours_ms": ev_median(lambda i: distCUDA2(pts), 20), "shape": "100k points"}
        if os.path.exists(knn_oracle.REF_LIB):
            knn_oracle.reference_dist_cuda2(pts); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                ref = knn_oracle.reference_dist_cuda2(pts)
            r["reference_simple_knn_ms"] = (time.perf_counter() - t0) / 5 * 1e3
            r["bit_identical_to_reference"] = bool(torch.equal(ref, distCUDA2(pts)))
        rows["f3_distCUDA2"] = r
    except Exception as e:  # noqa: BLE001
        rows["f3_distCUDA2"] = {"error": repr(e)[:200]}
    try:        # f4: extract_fields, reference defaults
        rawf = scene.to_raw_parameters(scene.make_cloud(100000, 0, seed=4, sigma=0.0128))
        tf = [torch.tensor(rawf[k], device=dev) for k in ("xyz", "opacity", "scaling", "rotation")]
        rows["f4_extract_fields"] = {"ours_ms": ev_median(lambda i: fields.extract_fields(*tf, resolution=128), 5), "shape": "100k Gaussians, 128^3, 16^3 blocks"}
    except Exception as e:  # noqa: BLE001
        rows["f4_extract_fields"] = {"error": repr(e)[:200]}
    return rows


def time_cpu_oracle(a, cloud, cams, ups, steps, warmup):
    """The CPU arm: the oracle port (float32, OpenMP over all host threads), one full view fwd+bwd per step."""
    from oracle import c_oracle
    c_oracle.set_threads(host_threads())            # torchrun exports OMP_NUM_THREADS=1: use every host core anyway
    bg = np.ones(3, np.float32)
    inputs = dict(means3D=cloud["means3D"], opacities=cloud["opacities"], shs=cloud["shs"], scales=cloud["scales"],
                  rotations=cloud["rotations"])
    times = []
    for i in range(warmup + steps):
        cam = cams[i % len(cams)]
        gC, gD, gA = ups[i % len(ups)]
        t0 = time.perf_counter()
        r = c_oracle.forward(**cam_settings(cam, a, bg), **inputs, dtype=np.float32)
        r.backward(gC, gD, gA)
        r.close()
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    sec = float(np.mean(times))
    return a.points / sec, sec, c_oracle.num_threads()


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(a, rank, world):
    if rank != 0:
        return
    cloud, cams, ups = build_scene(a)
    steps, warmup = max(1, min(a.steps, 5)), max(1, min(a.warmup, 1))
    val, sec, threads = time_cpu_oracle(a, cloud, cams, ups, steps, warmup)
    out = {
        "impl": "reference", "metric": "splats/sec fwd+bwd @ %dx%d" % (a.res, a.res), "value": val, "unit": "splats/s",
        "n_gpus": a.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "note": "CPU oracle port (oracle/dgr_oracle.c), the reference has no CPU path "
                   "and its CUDA op is not vendored (SURVEY.md §8c); steps capped at 5 to bound the run"},
        "cpu_baseline": {"value": val, "unit": "splats/s", "cores": threads, "kind": "port",
                         "sample": "%d full views fwd+bwd of the same workload" % steps, "cpu": cpu_model()},
        "e2e": {"value": val, "unit": "splats/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def run_ours(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from dreamgaussian_b200 import _lib, hostmem, multiview
    from dreamgaussian_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    cloud, cams, ups = build_scene(a)
    P, M, H, W = a.points, (a.sh_degree + 1) ** 2, a.res, a.res
    t = lambda x: torch.tensor(np.asarray(x, np.float32), device=dev)
    params = {k: t(v) for k, v in cloud.items()}
    bg = t(np.ones(3, np.float32))
    settings = [GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg, scale_modifier=1.0,
        viewmatrix=t(c.world_view_transform), projmatrix=t(c.full_proj_transform), sh_degree=a.sh_degree,
        campos=t(c.camera_center), prefiltered=False, debug=False) for c in cams]
    ups_d = [(t(u[0]), None, t(u[2])) for u in ups]
    vsr = multiview.ViewShardedRasterizer(P, M, dev)
    flush_buf = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)    # > 126 MB L2

    def step(i):
        v = (i * world + rank) % len(settings)                               # this rank's view of step i
        vsr.render_views(params, [settings[v]], [ups_d[v % len(ups_d)]])
        if world > 1:
            vsr.all_reduce()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # nvidia-smi samples every 20 ms; it is started before the warm-up so that even a short timed region is covered
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    for i in range(a.warmup):
        step(i)
    barrier()
    n_inst = 0
    # ---------------- timed region: K steps, device-resident inputs, L2 flushed between steps ----------------
    lib.dgr_reset_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    mark0 = sampler.mark()
    wall0 = time.perf_counter()
    for i in range(a.steps):
        flush_buf.zero_()
        evs[i][0].record()
        step(a.warmup + i)
        evs[i][1].record()
    barrier()
    wall1 = time.perf_counter()
    launches = int(lib.dgr_launch_count())
    clocks = sampler.stop((mark0, sampler.mark())) if rank == 0 else None
    ms_local = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    ms_t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_total = float(ms_t.item())
    ms_per_step = ms_total / a.steps
    value = P * world * a.steps / (ms_total * 1e-3)

    # ---------------- per-kernel CUDA-event timing (separate pass, same workload) ----------------
    kern = {}
    if rank == 0:
        lib.dgr_profile_enable(1)
        nprof = min(a.steps, 8)
        for i in range(nprof):
            flush_buf.zero_()
            vsr.render_views(params, [settings[i % len(settings)]], [ups_d[i % len(ups_d)]])
        torch.cuda.synchronize(dev)
        for name, ms in _lib.profile_collect():
            kern.setdefault(name, []).append(ms)
        lib.dgr_profile_enable(0)
        kern = {k: float(np.mean(v)) for k, v in kern.items()}
        # instance count of the bench views (host read-back the forward already does)
        _, _, _, _, st = __import__("dreamgaussian_b200.rasterizer", fromlist=["forward_impl"]).forward_impl(
            settings[0], params["means3D"], params["shs"], None, params["opacities"], params["scales"], params["rotations"], None)
        n_inst = int(st.num_rendered)
    if world > 1:
        barrier()

    # ---------------- end to end through the public API with HOST buffers ----------------
    # Every step copies ALL Gaussian inputs from pinned host memory to the device, runs GaussianRasterizer forward +
    # autograd backward, and copies the loss and ALL input gradients back to pinned host memory.  The three legs run on
    # three streams over a ring of 3 buffer sets, so step i's compute overlaps step i+1's upload and step i-1's download
    # (PCIe is full duplex); every byte of every step is moved inside the timed region.
    e2e = None
    if not a.no_e2e:
        names = ("means3D", "shs", "opacities", "scales", "rotations")
        # pinned buffers on the GPU's own NUMA node (dreamgaussian_b200/hostmem.py): cross-socket H2D runs at ~20 GB/s, local at ~53
        host = {k: hostmem.pinned_like(torch.tensor(cloud[k]), dev) for k in names}
        RING = 3
        dev_in = [{k: torch.empty_like(host[k], device=dev).requires_grad_(True) for k in names} for _ in range(RING)]
        grads_host = [{k: hostmem.pinned_empty(host[k].shape, host[k].dtype, dev) for k in names} for _ in range(RING)]
        loss_host = [hostmem.pinned_empty((1,), torch.float32, dev) for _ in range(RING)]
        h2d = sum(v.numel() * 4 for v in host.values())
        d2h = h2d + 4
        s_up, s_comp, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        ev_up = [torch.cuda.Event() for _ in range(RING)]
        ev_comp = [torch.cuda.Event() for _ in range(RING)]
        ev_down = [torch.cuda.Event() for _ in range(RING)]
        m2d = torch.zeros((P, 3), device=dev)

        def upload(i):
            r = i % RING
            with torch.cuda.stream(s_up):
                s_up.wait_event(ev_comp[r])                 # the previous user of this buffer set has finished computing
                with torch.no_grad():
                    for k in names:
                        dev_in[r][k].copy_(host[k], non_blocking=True)
                ev_up[r].record(s_up)

        def compute(i):
            r = i % RING
            v = (i * world + rank) % len(settings)
            with torch.cuda.stream(s_comp):
                s_comp.wait_event(ev_up[r])
                s_comp.wait_event(ev_down[r])               # its gradients of RING steps ago have been downloaded
                for k in names:
                    dev_in[r][k].grad = None
                color, radii, depth, alpha = GaussianRasterizer(raster_settings=settings[v])(
                    means3D=dev_in[r]["means3D"], means2D=m2d, shs=dev_in[r]["shs"], opacities=dev_in[r]["opacities"],
                    scales=dev_in[r]["scales"], rotations=dev_in[r]["rotations"])
                gC, _, gA = ups_d[v % len(ups_d)]
                loss = (color * gC).sum() + (alpha * gA).sum()
                loss.backward()
                if world > 1:
                    flat = torch.cat([dev_in[r][k].grad.reshape(-1) for k in names])
                    dist.all_reduce(flat)
                    o = 0
                    for k in names:
                        n = dev_in[r][k].grad.numel()
                        dev_in[r][k].grad.copy_(flat[o:o + n].view_as(dev_in[r][k].grad)); o += n
                ev_comp[r].record(s_comp)
                return loss.detach()

        def download(i, loss):
            r = i % RING
            with torch.cuda.stream(s_down):
                s_down.wait_event(ev_comp[r])
                for k in names:
                    g = dev_in[r][k].grad
                    g.record_stream(s_down)
                    grads_host[r][k].copy_(g, non_blocking=True)
                loss.record_stream(s_down)
                loss_host[r].copy_(loss.reshape(1), non_blocking=True)
                ev_down[r].record(s_down)

        def run(nsteps, first):
            upload(first)
            for i in range(first, first + nsteps):
                if i + 1 < first + nsteps:
                    upload(i + 1)
                ls = compute(i)
                download(i, ls)

        for r in range(RING):
            ev_comp[r].record(s_comp); ev_down[r].record(s_down)
        run(max(3, min(a.warmup, 6)), 0)
        barrier()
        e2e_steps = min(a.steps, 300)
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()                                          # default stream: ordered before the side streams by barrier()
        for st_ in (s_up, s_comp, s_down):
            st_.wait_event(e0)
        run(e2e_steps, 100)
        for st_ in (s_up, s_comp, s_down):
            torch.cuda.current_stream(dev).wait_stream(st_)
        e1.record()
        barrier()
        ms_e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
        e2e = {"value": P * world * e2e_steps / (float(ms_e.item()) * 1e-3), "unit": "splats/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": float(ms_e.item()) / e2e_steps, "steps": e2e_steps,
               "wall_ms_per_step": (time.perf_counter() - t0) * 1e3 / e2e_steps,
               "host_buffers_numa_local": bool(hostmem.gpu_local_cpus(dev)),
               "note": "public API (GaussianRasterizer + autograd); inputs from / gradients + loss to pinned host memory every step; "
                       "upload, compute and download of consecutive steps overlap on 3 streams"}

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    known = {k: v for k, v in kern.items() if kernel_algorithmic_bytes(k, P, M, H, W, n_inst)}
    dom = max(known, key=known.get) if known else None
    roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src}
    if dom:
        kb = kernel_algorithmic_bytes(dom, P, M, H, W, n_inst)
        ach = kb / (kern[dom] * 1e-3) / 1e9
        roof.update({"kernel": dom, "kernel_ms": kern[dom], "kernel_algorithmic_bytes": kb, "achieved": ach, "frac": ach / peak,
                     "kernels_ms": kern,
                     "kernels_frac": {k: (kernel_algorithmic_bytes(k, P, M, H, W, n_inst) or 0) / (v * 1e-3) / 1e9 / peak
                                      for k, v in kern.items()},
                     "step": {"algorithmic_bytes": algorithmic_bytes(P, M, H, W, n_inst), "n_inst": n_inst,
                              "achieved": algorithmic_bytes(P, M, H, W, n_inst) / (ms_per_step * 1e-3) / 1e9,
                              "frac": algorithmic_bytes(P, M, H, W, n_inst) / (ms_per_step * 1e-3) / 1e9 / peak},
                     "note": "render kernels are issue/MUFU-bound, not HBM-bound (DESIGN.md); fractions are against the HBM copy peak"})
    if dom:
        # DRAM traffic of the dominant kernel: from the committed `ncu --set full` capture of this same workload
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")))
            if tr.get("workload") == workload_name(a) and dom in tr["bytes_per_launch"]:
                roof["traffic"] = tr["bytes_per_launch"][dom]
                roof["traffic_source"] = tr["source"]
        except Exception:
            pass
    out = {
        "metric": "splats/sec fwd+bwd @ %dx%d" % (H, W), "value": value, "unit": "splats/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "views_per_gpu_per_step": 1, "view_set": a.views, "l2_flush_between_steps": True,
                   "n_inst_view0": n_inst, "parallelism": "view-sharded dp%d, replicated Gaussians, 1 all-reduce of %d MB per step (%s)"
                   % (world, vsr.grads.nbytes() >> 20, vsr.collective) if world > 1 else "single GPU",
                   "wall_ms_per_step_incl_flush": (wall1 - wall0) * 1e3 / a.steps},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "e2e": e2e,
    }
    if world == 1 and not a.no_cpu_baseline:
        val, sec, threads = time_cpu_oracle(a, cloud, cams, ups, 2, 1)
        out["cpu_baseline"] = {"value": val, "unit": "splats/s", "cores": threads, "kind": "port",
                               "sample": "2 full views fwd+bwd of the same workload (oracle/dgr_oracle.c, float32, OpenMP)",
                               "seconds_per_view": sec, "cpu": cpu_model()}
    if world == 1 and not a.no_rows:
        out["next_rows"] = measure_next_rows(dev)
    print(json.dumps(out), flush=True)


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        run_ours(a, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
