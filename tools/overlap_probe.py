"""How much idle issue capacity do the render kernels leave?  Two independent backward (or forward) passes of two views
run first back-to-back on one stream, then concurrently on two streams: if the pair takes much less than 2x one pass, the
kernel is parallelism-starved (590 non-empty tiles x 4 warps = 16 warps per SM at the bench shape), not throughput-bound."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from dreamgaussian_b200 import scene, rasterizer as R
from dreamgaussian_b200.rasterizer import GaussianRasterizationSettings

dev = torch.device("cuda", 0)
P, res = 100000, 800
cloud = scene.make_cloud(P, 3, seed=0, opacity="trained", anisotropic=True)
cams = scene.bench_views(8, res, res)
t = lambda x: torch.tensor(np.asarray(x, np.float32), device=dev)
params = {k: t(v) for k, v in cloud.items()}
bg = t(np.ones(3, np.float32))
settings = [GaussianRasterizationSettings(image_height=res, image_width=res, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg,
            scale_modifier=1.0, viewmatrix=t(c.world_view_transform), projmatrix=t(c.full_proj_transform), sh_degree=3,
            campos=t(c.camera_center), prefiltered=False, debug=False) for c in cams]
rng = np.random.default_rng(17)
gC, gA = t(rng.normal(size=(3, res, res))), t(rng.normal(size=(1, res, res)))
f32 = dict(dtype=torch.float32, device=dev)
def gbuf():
    return [torch.empty((P, 3), **f32), torch.empty((P, 3), **f32), torch.empty((P, 16, 3), **f32), None, torch.empty((P, 1), **f32),
            torch.empty((P, 3), **f32), torch.empty((P, 4), **f32), None]
states = []
for v in (0, 3):
    out = R.forward_impl(settings[v], params["means3D"], params["shs"], None, params["opacities"], params["scales"], params["rotations"], None)
    states.append(out[-1])
bufs = [gbuf(), gbuf()]
s = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
def bwd(i):
    R.backward_impl(states[i], gC, None, gA, *bufs[i])
def fwd(i):
    R.forward_impl(settings[(0, 3)[i]], params["means3D"], params["shs"], None, params["opacities"], params["scales"], params["rotations"], None)
def timed(fn, reps=30):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(e0); 
        for st in s: torch.cuda.current_stream().wait_stream(st)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))
for name, op in (("backward", bwd), ("forward", fwd)):
    def serial(e0):
        op(0); op(1)
    def single(e0):
        op(0)
    def conc(e0):
        for i in (0, 1):
            s[i].wait_event(e0)
            with torch.cuda.stream(s[i]): op(i)
    for _ in range(3): serial(None)
    a, b, c = timed(single), timed(serial), timed(conc)
    print("%s: one view %.1f us, two back-to-back %.1f us, two concurrent %.1f us  (concurrent / serial = %.2f)" % (name, a, b, c, c / b))
