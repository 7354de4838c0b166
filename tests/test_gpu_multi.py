"""GPU: the ray-sharded near field on 2 GPUs (NCCL all-reduce of the bin planes) equals the
single-GPU result.  Skipped on a 1-GPU box; the sharding/reduce logic itself is covered on CPU
by tests/test_host_logic.py (gloo, world size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from sosat_optics_b200 import dist as sdist, opt_analyze, ot_geo
        t = ot_geo.SatGeo()
        t.n_scan = 700
        t.y_source = ot_geo.y_lyot + 300
        t.lambda_ = opt_analyze.ghz_to_m(90)
        t.k = 2 * np.pi / t.lambda_
        b, cnt, stats = sdist.near_field_binned_distributed(t, [[0, 0, 0], [100, 0, 60]], 128, 42.0)
        q.put((rank, b.cpu().numpy(), cnt.cpu().numpy(), stats))
    finally:
        dist.destroy_process_group()


def test_two_gpu_ray_sharding_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from sosat_optics_b200 import opt_analyze, ot_geo, ray_trace
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    t = ot_geo.SatGeo()
    t.n_scan = 700
    t.y_source = ot_geo.y_lyot + 300
    t.lambda_ = opt_analyze.ghz_to_m(90)
    t.k = 2 * np.pi / t.lambda_
    one = ray_trace.near_field_binned(t, [[0, 0, 0], [100, 0, 60]], grid_n=128, extent_cm=42.0)
    for rank, b, cnt, stats in res:
        assert np.array_equal(cnt, one.count)
        assert np.abs(b - one.b).max() < 1e-4 * np.abs(one.b).max()
        assert np.array_equal(stats[:, [0, 7]], one.stats[:, [0, 7]])
        assert np.allclose(stats[:, 1], one.stats[:, 1], rtol=1e-12)
