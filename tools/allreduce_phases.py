"""torchrun diagnostic: where the multi-GPU step of the bench workload spends its time.

Per step, CUDA events on the rank's stream around: render_views | barrier | all-reduce kernel | barrier, for the library's own
collective (host barriers), the in-kernel-barrier variant and NCCL; plus the per-view spread of the single-GPU step (at N ranks
the step waits for the slowest view).  Prints medians, max over ranks.
"""
import argparse, ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from dreamgaussian_b200 import _lib, multiview, scene
from dreamgaussian_b200.rasterizer import GaussianRasterizationSettings

ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=120)
ap.add_argument("--modes", default="", help="comma-separated subset of the modes below"); a = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world)
lib = _lib.load()
P, res, deg = 100000, 800, 3
cloud = scene.make_cloud(P, deg, seed=0, opacity="trained", anisotropic=True)
cams = scene.bench_views(8, res, res)
t = lambda x: torch.tensor(np.asarray(x, np.float32), device=dev)
params = {k: t(v) for k, v in cloud.items()}
bg = t(np.ones(3, np.float32))
settings = [GaussianRasterizationSettings(image_height=res, image_width=res, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg,
            scale_modifier=1.0, viewmatrix=t(c.world_view_transform), projmatrix=t(c.full_proj_transform), sh_degree=deg,
            campos=t(c.camera_center), prefiltered=False, debug=False) for c in cams]
rng = np.random.default_rng(17)
up = (t(rng.normal(size=(3, res, res))), None, t(rng.normal(size=(1, res, res))))
flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)


def med_max(x):
    v = torch.tensor([float(np.median(x))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return round(float(v) * 1e3, 1)


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


out = {"world": world}
# (1) per-view spread of the local step (no collective)
vsr = multiview.ViewShardedRasterizer(P, (deg + 1) ** 2, dev, peer_allreduce=False)
per_view = []
for v in range(8):
    ts = []
    for i in range(12):
        flush.zero_(); e0 = ev(); vsr.render_views(params, [settings[v]], [up]); e1 = ev(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    per_view.append(round(float(np.median(ts[2:])) * 1e3, 1))
out["per_view_us"] = per_view
del vsr

if world > 1:
    for mode, env in (("allreduce_inkernel", {}), ("allreduce_hostbar", {"DGR_INKERNEL_BARRIERS": "0"}), ("allreduce_p2p", {"DGR_NO_MULTIMEM": "1"}),
                      ("push_inkernel", {"DGR_PUSH": "1"}), ("push_hostbar", {"DGR_PUSH": "1", "DGR_INKERNEL_BARRIERS": "0"}),
                      ("multimem_forced", {"DGR_FORCE_MULTIMEM": "1"}), ("nccl", {"DGR_NO_PEER": "1"})):
        if a.modes and mode not in a.modes.split(","):
            continue
        for k in ("DGR_INKERNEL_BARRIERS", "DGR_NO_MULTIMEM", "DGR_NO_PEER", "DGR_PUSH", "DGR_FORCE_MULTIMEM"):
            os.environ.pop(k, None)
        os.environ.update(env)
        vsr = multiview.ViewShardedRasterizer(P, (deg + 1) ** 2, dev)
        own = vsr._hdl is not None
        rows = []
        for i in range(a.steps + 5):
            v = (i * world + rank) % 8
            flush.zero_()
            e0 = ev(); vsr.render_views(params, [settings[v]], [up]); e1 = ev()
            if own and not vsr._inkernel and not vsr._use_push:
                vsr._epoch += 1
                vsr._hdl.barrier(channel=0); e2 = ev()
                st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                _lib.check(lib.dgr_peer_allreduce(ctypes.cast(vsr._ptrs, ctypes.c_void_p), world, rank, ctypes.c_uint64(vsr.grads.padded),
                                                  ctypes.c_uint64(vsr._mc), None, ctypes.c_uint32(vsr._epoch), st)); e3 = ev()
                vsr._hdl.barrier(channel=1); e4 = ev()
            else:
                vsr.all_reduce(); e2 = e3 = e4 = ev()
            rows.append((e0, e1, e2, e3, e4))
        torch.cuda.synchronize()
        rows = rows[5:]
        ph = lambda i, j: [r[i].elapsed_time(r[j]) for r in rows]
        out[mode] = {"collective": vsr.collective, "render_us": med_max(ph(0, 1)), "barrier0_us": med_max(ph(1, 2)), "kernel_us": med_max(ph(2, 3)),
                     "barrier1_us": med_max(ph(3, 4)), "step_us": med_max(ph(0, 4)),
                     "mean_step_us": round(float(np.mean(ph(0, 4))) * 1e3, 1)}
        del vsr
        dist.barrier()
if rank == 0:
    print(json.dumps(out, indent=1))
if world > 1:
    dist.destroy_process_group()
