"""GPU: the ray-sharded near field on 2 GPUs (NCCL all-reduce of the bin planes, and the fused
reduce over peer-mapped memory) equals the single-GPU result.  Skipped on a 1-GPU box; the sharding/reduce logic itself is covered on CPU
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


RX_STEPS = ([[0, 0, 0], [100, 0, 60]], [[-80, 0, 40], [30, 0, -120]])


def _peer_worker(rank, world, port, q):
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
        peer = sdist.PeerPlanes(2, 1, 128)
        out = []
        for rx in RX_STEPS:  # the same buffers twice: a stale read of a peer's planes would show
            b, _, stats = sdist.near_field_binned_distributed(t, rx, 128, 42.0, peer=peer)
            out.append((b.cpu().numpy(), stats))
        # far field of the first field on rank 0, written where the peers can read it, then shipped
        # to ONE shared host array by both ranks (each its slab, over its own PCIe link)
        from sosat_optics_b200 import opt_analyze as oa
        host = sdist.SharedHostArray((2, 1, 128, 128), np.complex64)
        if rank == 0:
            peer.out.zero_()
            oa.far_field(peer.field[0, 0], device=True, out=peer.out[0, 0])
            oa.far_field(peer.field[1, 0], device=True, out=peer.out[1, 0])
        peer.gather_to_host(host)
        shipped = host.view().copy()
        want = peer.out.cpu().numpy() if rank == 0 else None
        torch.cuda.synchronize()
        dist.barrier()
        host.close()
        peer.close()
        q.put((rank, out, shipped, want))
    finally:
        dist.destroy_process_group()


def test_two_gpu_peer_memory_reduce_matches_single_gpu():
    """csrc/peer.cu: every rank reduces its slab of cells from both ranks' planes over peer
    memory, normalises it and stores it to both ranks."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from sosat_optics_b200 import opt_analyze, ot_geo, ray_trace
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    t = ot_geo.SatGeo()
    t.n_scan = 700
    t.y_source = ot_geo.y_lyot + 300
    t.lambda_ = opt_analyze.ghz_to_m(90)
    t.k = 2 * np.pi / t.lambda_
    for step, rx in enumerate(RX_STEPS):
        one = ray_trace.near_field_binned(t, rx, grid_n=128, extent_cm=42.0)
        for rank, out, _, _ in res:
            b, stats = out[step]
            assert np.abs(b - one.b).max() < 1e-4 * np.abs(one.b).max(), (step, rank)
            assert np.array_equal(stats[:, [0, 7]], one.stats[:, [0, 7]])
            assert np.allclose(stats[:, 1], one.stats[:, 1], rtol=1e-12)
    # gather_to_host: both ranks see rank 0's far fields whole in the shared host array
    want = res[0][3]
    assert want is not None and np.abs(want).max() > 0
    for rank, _, shipped, _ in res:
        assert np.array_equal(shipped, want), rank


def _sweep_setup():
    from sosat_optics_b200 import geometry, opt_analyze, ot_geo
    t = ot_geo.SatGeo()
    t.n_scan = 300
    t.y_source = ot_geo.y_lyot + 300
    t.lambda_ = opt_analyze.ghz_to_m(90)
    t.k = 2 * np.pi / t.lambda_
    table = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                 "feedhorn_fwhm.npz"))["table"]
    specs = geometry.freq_specs_from_table([90.0, 150.0], table)
    g = np.linspace(-120, 120, 3)
    rxs = np.array([[x, 0.0, z] for x in g for z in g][:7])   # 7 receivers: an uneven deal
    return t, rxs, specs


def _sweep_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from sosat_optics_b200 import dist as sdist
        t, rxs, specs = _sweep_setup()
        res = sdist.beam_sweep_distributed(t, rxs, specs, grid_n=128, extent_cm=42.0, pad=64)
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_two_gpu_receiver_sweep_matches_single_gpu():
    """BASELINE configs[3] sharded over 2 GPUs (receivers round-robin, scalars gathered) gives,
    on every rank, exactly what the single-GPU beam_sweep gives for the whole receiver list."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from sosat_optics_b200 import ray_trace
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sweep_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    t, rxs, specs = _sweep_setup()
    one = ray_trace.beam_sweep(t, rxs, specs, grid_n=128, extent_cm=42.0, pad=64)
    assert one["fwhm_arcmin"].shape == (7, 2) and (one["fwhm_arcmin"] > 0).all()
    for rank, got in res:
        assert set(got) == set(one)
        for k in ("n_valid", "n_binned", "peak_index"):
            assert np.array_equal(got[k], one[k]), (rank, k)
        assert np.allclose(got["mean_opl"], one["mean_opl"], rtol=1e-13)
        # the binning's fp32 atomics are order dependent: far-field scalars agree to that level
        assert np.allclose(got["peak"], one["peak"], rtol=1e-4)
        lam = np.array([2 * np.pi / float(k) for k, _, _ in specs])
        cell = lam / (256 * (42.0 / 128 / 100.0)) * (180 * 60 / np.pi)   # arcmin per far-field cell
        assert (np.abs(got["fwhm_arcmin"] - one["fwhm_arcmin"]) <= 1.001 * cell[None, :]).all()
        assert np.mean(got["fwhm_arcmin"] == one["fwhm_arcmin"]) >= 0.8


def test_side_stream_and_second_device_in_one_process():
    """The library launches on the caller's current stream and keeps per-device state
    (geometry in __constant__ memory, DFT plans): the same calls on a side stream, and on
    cuda:1 of the same process when there is one, give the default-stream / cuda:0 results."""
    import torch
    from sosat_optics_b200 import opt_analyze, ot_geo, ray_trace
    t = ot_geo.SatGeo()
    t.n_scan = 96
    t.y_source = ot_geo.y_lyot + 300
    t.lambda_ = opt_analyze.ghz_to_m(90)
    t.k = 2 * np.pi / t.lambda_
    rng = np.random.default_rng(11)
    b = (rng.standard_normal((200, 200)) + 1j * rng.standard_normal((200, 200))).astype(np.complex64)

    def run():
        bf = ray_trace.near_field_binned(t, [25, 0, -40], grid_n=48, extent_cm=42.0)
        return bf.b.copy(), bf.count.copy(), opt_analyze.far_field(b), ray_trace.rx_to_lyot([25, 0, -40], t)

    torch.cuda.set_device(0)
    base = run()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        on_side = run()
    side.synchronize()
    devices = [("side stream", on_side)]
    if torch.cuda.device_count() >= 2:
        torch.cuda.set_device(1)
        try:
            devices.append(("cuda:1", run()))
        finally:
            torch.cuda.set_device(0)
    for name, got in devices:
        assert np.array_equal(got[1], base[1]), name
        assert np.abs(got[0] - base[0]).max() <= 1e-5 * np.abs(base[0]).max(), name
        assert np.abs(got[2] - base[2]).max() <= 1e-6 * np.abs(base[2]).max(), name
        assert np.array_equal(np.nan_to_num(got[3]), np.nan_to_num(base[3])), name
    again = run()  # and device 0 still has its own geometry / plans
    assert np.array_equal(again[1], base[1])
