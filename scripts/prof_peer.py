"""Time the fused peer-memory reduce kernel alone (no barriers): run under torchrun, N ranks.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/prof_peer.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from sosat_optics_b200 import _lib, dist as sdist  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
N = 2048
peer = sdist.PeerPlanes(1, 1, N)
peer.cnt.fill_(1.0)
peer.re.fill_(float(rank))
peer.stats[0, 0] = 1.0
lib, L = peer.lib, peer.layout
b0, b1 = sdist.slab_cells(N * N, world, rank)
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for dest in ([0], list(range(world))):
    c_out = (ctypes.c_void_p * len(dest))(*[peer.peer_base[d] for d in dest])
    ts = []
    for rep in range(12):
        peer.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.sosat_bins_reduce_field(peer._c_peers, world, L["re"], L["im"], L["cnt"], L["stats"],
                                               b0, b1, 1.0, 0.0, c_out, len(dest), L["field"], None, stream))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        peer.barrier()
    e1.record()
    torch.cuda.synchronize()
    print("rank %d dest %s: kernel best %.3f ms median %.3f ms; barrier %.3f ms; field[0,0,5,5]=%s" % (
        rank, dest, min(ts[2:]), sorted(ts[2:])[len(ts[2:]) // 2], e0.elapsed_time(e1) / 10,
        peer.field[0, 0, 5, 5].item()), flush=True)
dist.barrier()
peer.close()
dist.destroy_process_group()
