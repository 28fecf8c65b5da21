"""torchrun diagnostic: the library's NVLink all-reduce vs NCCL on the flat gradient buffer (correctness + time)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from dreamgaussian_b200 import multiview
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", rank=rank, world_size=world)
P, M = 100000, 16
for mode in ("auto", "p2p", "nccl"):
    os.environ["DGR_NO_MULTIMEM"] = "1" if mode == "p2p" else "0"
    vsr = multiview.ViewShardedRasterizer(P, M, dev, peer_allreduce=(mode != "nccl"))
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    src = torch.randn(vsr.grads.flat.numel(), device=dev, generator=g)
    ref = src.clone(); dist.all_reduce(ref)
    vsr.grads.flat.copy_(src)
    out = vsr.all_reduce().clone()
    err = float((out - ref).abs().max()) / float(ref.abs().max())
    # identical on every rank?
    chk = out.double().sum().reshape(1).clone(); lst = [torch.zeros_like(chk) for _ in range(world)]; dist.all_gather(lst, chk)
    same = all(float(x) == float(lst[0]) for x in lst)
    evs = []
    for it in range(30):
        vsr.grads.flat.copy_(src); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); vsr.all_reduce(); e1.record(); torch.cuda.synchronize(); evs.append(e0.elapsed_time(e1))
    t = torch.tensor([sorted(evs)[len(evs) // 2]], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("mode %-5s -> %-28s  rel.err vs NCCL %.2e  identical on ranks %s  median %.1f us  (%.1f MB)" %
              (mode, vsr.collective, err, same, float(t) * 1e3, vsr.grads.nbytes() / 1e6), flush=True)
    if mode != "nccl" and vsr._hdl is None and rank == 0:
        print("   fallback reason:", getattr(vsr, "_why_nccl", None))
    del vsr
dist.destroy_process_group()
