"""Where PeerPlanes.gather_to_host spends its time at N ranks (torchrun):
torchrun --nproc-per-node N scripts/prof_gather.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from sosat_optics_b200 import dist as sdist

local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n = 2048
peer = sdist.PeerPlanes(1, 1, n)
host = sdist.SharedHostArray((1, 1, n, n), np.complex64)
if rank == 0:
    peer.out.copy_(torch.randn(1, 1, n, n, dtype=torch.complex64, device="cuda"))
torch.cuda.synchronize(); dist.barrier()
ev = lambda: torch.cuda.Event(enable_timing=True)
total = n * n
b0, b1 = sdist.slab_cells(total, world, rank)
for rep in range(6):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    peer.gather_to_host(host)
    t_all = (time.perf_counter() - t0) * 1e3
    # the pieces, one after the other
    e = [ev() for _ in range(5)]
    torch.cuda.synchronize(); dist.barrier()
    e[0].record(); peer.barrier(); e[1].record()
    peer._stage.copy_(peer._peer_out[b0:b1]); e[2].record()
    host.tensor[b0:b1].copy_(peer._stage, non_blocking=True); e[3].record()
    peer.barrier(); e[4].record()
    torch.cuda.synchronize()
    parts = [e[i].elapsed_time(e[i + 1]) for i in range(4)]
    t = torch.tensor([t_all] + parts, dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0 and rep >= 2:
        print("world %d gather_to_host %.3f ms | barrier %.3f  peer->local %.3f  d2h %.3f  barrier %.3f (max over ranks)"
              % ((world,) + tuple(t.tolist())), flush=True)
# one rank alone copying everything, for comparison
if rank == 0:
    full = torch.empty(total, dtype=torch.complex64, device="cuda")
    pin = torch.empty(total, dtype=torch.complex64).pin_memory()
    for _ in range(3):
        a, b = ev(), ev(); a.record(); pin.copy_(full, non_blocking=True); b.record(); torch.cuda.synchronize()
    print("rank 0 alone, 33.5 MB d2h: %.3f ms" % a.elapsed_time(b), flush=True)
    part = full[:total // world]
    for _ in range(3):
        a, b = ev(), ev(); a.record(); pin[:total // world].copy_(part, non_blocking=True); b.record(); torch.cuda.synchronize()
    print("rank 0 alone, 1/%d slab d2h (pinned torch buffer): %.3f ms" % (world, a.elapsed_time(b)), flush=True)
    for _ in range(3):
        a, b = ev(), ev(); a.record(); host.tensor[b0:b1].copy_(part, non_blocking=True); b.record(); torch.cuda.synchronize()
    print("rank 0 alone, 1/%d slab d2h (shared registered array): %.3f ms" % (world, a.elapsed_time(b)), flush=True)
dist.barrier()
host.close(); peer.close()
dist.destroy_process_group()
