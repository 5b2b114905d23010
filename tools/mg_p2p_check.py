"""Multi-GPU check of the peer-memory halo path (run under torchrun, one rank per GPU):
PeerRing (halos read from the neighbours' HBM inside the kernel) must give bit-identical results to the
all_gather exchange, for periodic and non-periodic boundaries."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import savgol_b200 as sg
from savgol_b200 import dist as sgd

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
ok = True
for mode in ("periodic", "reflect", "polynomial", "constant"):
    for n, L in ((32, 1 << 22), (5, 100_003 + 17 * rank)):
        g = torch.Generator(device="cuda"); g.manual_seed(100 + rank)
        x = torch.randn(L, device="cuda", generator=g)
        f = sg.SavgolFilter(n, 4, 2, 1.0, mode)
        y_ref = sgd.apply_partitioned(f, x)
        ring = sgd.PeerRing(x, n, mode == "periodic")
        torch.cuda.synchronize(); dist.barrier()
        y = ring.apply(f, torch.empty_like(x))
        torch.cuda.synchronize(); dist.barrier()
        same = bool(torch.equal(y, y_ref))
        ok &= same
        ring.close()
        if not same:
            print(f"rank {rank} mode {mode} n {n}: MISMATCH max {float((y - y_ref).abs().max())}")
# row-band sharding of one image: every rank filters its band (halo rows by all_gather), result == the same
# rows of the whole-image filter computed locally
gi = torch.Generator(device="cuda"); gi.manual_seed(4242)
img = torch.rand(1500, 1024, device="cuda", generator=gi)
f2 = sg.Savgol2DFilter(7, 7, 3)
for boundary in ("constant", "reflect"):
    a, b = sgd.shard_range(img.shape[0], rank, world)
    yb = sgd.apply_image_bands(f2, img[a:b].contiguous(), boundary)
    same = bool(torch.equal(yb, f2.apply(img, boundary)[a:b]))
    ok &= same
    if not same:
        print(f"rank {rank} band {boundary}: MISMATCH")
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("(1D peer-memory halos + 2D image bands)", end=" ")
    print("p2p halo check:", "PASS" if int(t.item()) == 1 else "FAIL", "world", world)
dist.destroy_process_group()
