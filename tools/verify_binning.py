"""Diagnostic: check the binning buffers of one forward (ranges, sorted ids/records) for consistency, twice."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from dreamgaussian_b200 import scene, rasterizer as R, _lib
dev = torch.device("cuda")
CFG = eval(os.environ.get("VB_CFG", "((1000000, 1600, 0.006),)"))
def al(v, a=256): return (v + a - 1) // a * a
for (P, res, sigma) in CFG:
    deg = 3
    cloud = scene.make_cloud(P, deg, seed=2, sigma=sigma)
    params = {k: torch.tensor(v, device=dev) for k, v in cloud.items()}
    cam = scene.orbit_camera(10, 30, 2.0, res, res)
    t = lambda a: torch.tensor(np.asarray(a, np.float32), device=dev)
    rs = R.GaussianRasterizationSettings(image_height=res, image_width=res, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=t(np.zeros(3)),
        scale_modifier=1.0, viewmatrix=t(cam.world_view_transform), projmatrix=t(cam.full_proj_transform), sh_degree=deg,
        campos=t(cam.camera_center), prefiltered=False, debug=False)
    outs = []
    for rep in range(3):
        c, r, d, a, st = R.forward_impl(rs, params["means3D"], params["shs"], None, params["opacities"], params["scales"], params["rotations"], None)
        torch.cuda.synchronize()
        tiles = ((res + 15) // 16) ** 2
        img = st.image.cpu().numpy(); binb = st.binning.cpu().numpy()
        ranges = img[: tiles * 8].view(np.uint32).reshape(tiles, 2)
        cap = st.capacity
        off_ids = al(cap * 8); off_rec = off_ids + al(cap * 4)
        ids = binb[off_ids: off_ids + cap * 4].view(np.uint32)
        rec = binb[off_rec: off_rec + cap * 48].view(np.float32).reshape(cap, 12)
        n = st.num_rendered
        cnt = (ranges[:, 1] - ranges[:, 0]).astype(np.int64)
        ok_sum = int(cnt.sum()) == n and (ranges[1:, 0] == ranges[:-1, 1]).all() and ranges[0, 0] == 0
        bad_id = int((ids[:n] >= P).sum())
        depth = rec[:n, 6]
        # sortedness inside every tile
        tile_of = np.repeat(np.arange(tiles), cnt)
        same = tile_of[1:] == tile_of[:-1]
        unsorted = int(((depth[1:] < depth[:-1]) & same).sum())
        ties_bad = int(((depth[1:] == depth[:-1]) & same & (ids[1:n] <= ids[:n - 1])).sum())
        print(P, res, "rep", rep, "n_inst", n, "cap", cap, "max/tile", int(cnt.max()), "big tiles", int((cnt > 4096).sum()), "sum ok", bool(ok_sum),
              "bad ids", bad_id, "unsorted pairs", unsorted, "bad ties", ties_bad, "hint", R._CAPACITY_HINT.get((0, P, res, res)), flush=True)
        outs.append((c.clone(), ids[:n].copy(), ranges.copy()))
    print("   fwd identical across reps:", torch.equal(outs[0][0], outs[1][0]), torch.equal(outs[0][0], outs[2][0]),
          "ids identical:", (outs[0][1] == outs[1][1]).all(), (outs[0][1] == outs[2][1]).all())
