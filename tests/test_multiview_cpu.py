"""CPU (gloo, world_size 2) tests of the host-side multi-GPU logic: view sharding, the flat gradient layout and the
all-reduce of per-Gaussian gradients.  The kernels themselves need a GPU; here each rank fills its flat
buffer with the float64 CPU oracle's gradients of ITS views and the reduced buffer must equal the sum over all views."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as h
from dreamgaussian_b200 import multiview, scene


def test_shard_views_partitions_every_view_once():
    for V in (1, 7, 8, 64):
        for G in (1, 2, 3, 8):
            got = sorted(v for r in range(G) for v in multiview.shard_views(V, r, G))
            assert got == list(range(V))
            sizes = [len(multiview.shard_views(V, r, G)) for r in range(G)]
            assert max(sizes) - min(sizes) <= 1


def test_flat_grads_layout_is_one_contiguous_buffer():
    P, M = 10, 4
    fg = multiview.FlatGrads(P, M, "cpu")
    assert fg.flat.numel() == P * (3 + 3 * M + 1 + 3 + 4 + 3) and fg.flat.is_contiguous()
    o = 0
    for name, shape in (("means3D", (P, 3)), ("shs", (P, M, 3)), ("opacities", (P, 1)), ("scales", (P, 3)), ("rotations", (P, 4)),
                        ("means2D", (P, 3))):
        v = fg.views[name]
        assert tuple(v.shape) == shape and v.data_ptr() == fg.flat.data_ptr() + 4 * o
        o += v.numel()
    fg.views["scales"].fill_(2.0)
    assert fg.flat.sum().item() == 2.0 * 3 * P


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, P, deg, res, n_views, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cloud = scene.make_cloud(P, deg, seed=0, sigma=0.06)
        cams = scene.bench_views(n_views, res, res)
        vsr = multiview.ViewShardedRasterizer(P, (deg + 1) ** 2, "cpu")
        inputs = dict(means3D=cloud["means3D"], opacities=cloud["opacities"], shs=cloud["shs"], scales=cloud["scales"],
                      rotations=cloud["rotations"])
        for v in multiview.shard_views(n_views, rank, world):
            cam = cams[v]
            s = dict(image_height=res, image_width=res, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=np.ones(3), scale_modifier=1.0,
                     viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, sh_degree=deg, campos=cam.camera_center)
            g = h.upstream_grads(res, res, seed=100 + v)
            ref = h.run_oracle(s, inputs, g)
            for k, view in vsr.grads.views.items():          # what the CUDA backward does with accumulate=1
                view += torch.tensor(ref["grads"][k].astype(np.float32)).view_as(view)
        vsr.pg = None
        flat = vsr.all_reduce().clone()
        if rank == 0:
            np.save(os.path.join(out_dir, "flat.npy"), flat.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_all_reduce_sums_view_gradients(tmp_path):
    P, deg, res, n_views = 300, 1, 32, 4
    port = _free_port()
    mp.spawn(_worker, args=(2, port, P, deg, res, n_views, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "flat.npy"))
    # single-process reference: all views summed
    cloud = scene.make_cloud(P, deg, seed=0, sigma=0.06)
    cams = scene.bench_views(n_views, res, res)
    inputs = dict(means3D=cloud["means3D"], opacities=cloud["opacities"], shs=cloud["shs"], scales=cloud["scales"],
                  rotations=cloud["rotations"])
    fg = multiview.FlatGrads(P, (deg + 1) ** 2, "cpu")
    for v in range(n_views):
        cam = cams[v]
        s = dict(image_height=res, image_width=res, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=np.ones(3), scale_modifier=1.0,
                 viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, sh_degree=deg, campos=cam.camera_center)
        ref = h.run_oracle(s, inputs, h.upstream_grads(res, res, seed=100 + v))
        for k, view in fg.views.items():
            view += torch.tensor(ref["grads"][k].astype(np.float32)).view_as(view)
    np.testing.assert_allclose(got, fg.flat.numpy(), rtol=1e-5, atol=1e-5 * np.abs(fg.flat.numpy()).max())


def _worker_empty_shard(rank, world, port, out_dir):
    """a rank with no view this iteration must contribute zeros, not last iteration's reduced sum."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P, M = 50, 4
        vsr = multiview.ViewShardedRasterizer(P, M, "cpu")
        results = []
        for it in range(2):                                     # two iterations with ONE view for two ranks
            mine = multiview.shard_views(1, rank, world)
            if mine:
                vsr.grads.flat.fill_(1.0 + it)                    # stands for render_views() overwriting with this view's gradient
            else:
                vsr.render_views({}, [], [])                    # no local view
            results.append(vsr.all_reduce().clone())
        torch.save(results, os.path.join(out_dir, "r%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def test_rank_without_a_view_contributes_zero(tmp_path):
    port = _free_port()
    mp.spawn(_worker_empty_shard, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        res = torch.load(os.path.join(str(tmp_path), "r%d.pt" % r))
        assert float(res[0].min()) == 1.0 and float(res[0].max()) == 1.0
        assert float(res[1].min()) == 2.0 and float(res[1].max()) == 2.0      # would be 3.0 with the stale buffer
