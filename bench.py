#!/usr/bin/env python
"""Benchmark of the beam-simulation hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the whole hot path over BASELINE.json configs[4] ("high-resolution
stress"): 1.000014e9 rays (n_scan = 31623) traced feed -> source plane from a boresight feed
at 90 GHz, binned into the 2048^2 complex near field (stop diameter 42 cm), bins all-reduced
over NCCL when N > 1 (rays are sharded by launch-row blocks: strong scaling), then normalised
and transformed to the 2048^2 far field.  The ray bundle is a deterministic grid, so the
"synthetic data" is the reference's own launch grid at a larger n_scan.

`value` = launched rays per second (the unit of the reference's tqdm it/s) with everything
resident on the device; `e2e` = the same through the public host-buffer API (parameters in,
2048^2 complex64 far field out to pinned host memory, copies inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_SCAN = 31623            # 1.000014e9 rays (cfg 5)
GRID_N = 2048
EXTENT_CM = 42.0
GHZ = 90.0
FLOP_PER_RAY = 1300.0     # algorithmic work per launched ray
METRIC = "rays/sec traced feed->window"
WORKLOAD = ("configs[4] high-resolution stress: 1.000014e9 rays/feed (n_scan=31623), boresight, "
            "90 GHz, 2048^2 near-field grid -> 2048^2 far field")


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def _tele_geo(ot_geo, opt_analyze, table, n_scan):
    t = ot_geo.SatGeo()
    t.n_scan = n_scan
    t.y_source = ot_geo.y_lyot + 300
    t.lambda_ = opt_analyze.ghz_to_m(GHZ)
    t.k = 2 * np.pi / t.lambda_
    row = table[table[:, 0] == int(GHZ)][0]
    t.th_fwhp_x, t.th_fwhp_y = np.deg2rad(row[1]), np.deg2rad(row[2])
    return t


def _feed_table():
    return np.load(os.path.join(ROOT, "feedhorn", "fwhm_feedhorns.npy"))


class ClockSampler(threading.Thread):
    """Samples SM clock and clock-event reasons of one GPU during the timed region (NVML)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake", 0x2: "applications_clocks"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def sample(self):
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            try:
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        while self.ok and not self._stop_evt.is_set():
            self.sample()
            time.sleep(0.005)  # short timed regions (8 GPUs: 40 ms) still get several samples

    def result(self):
        self._stop_evt.set()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
def shared_config(n_scan):
    """`config` of the JSON line: identical in the GPU arm and the --impl reference arm."""
    return {"workload": WORKLOAD if n_scan == N_SCAN else f"REDUCED n_scan={n_scan}: " + WORKLOAD,
            "n_scan": n_scan, "rays_per_step": n_scan * n_scan, "grid_n": GRID_N,
            "extent_cm": EXTENT_CM, "ghz": GHZ}


def cpu_reference_sample(n_scan_sample, threads=0, slab_rays=4_000_000):
    """The reference's CPU path on a bounded sample of the workload, restated in oracle/ (the
    reference itself is Python and cannot travel to the GPU box): C trace with the bracketed
    Brent solve on all host threads (slabs of launch rows, so the 136 B/ray table stays small),
    numpy binning (bincount), numpy FFT far field.  The per-ray part (trace + binning) and the
    fixed part (bins -> field, 2048^2 FFT) are timed separately so that the fixed cost can be
    amortised over the full bundle exactly as the GPU step amortises it.
    Returns a dict: rays, cores, t_trace, t_bin, t_fixed, t_total [s]."""
    from oracle import oracle
    t = oracle.default_tele_geo(n_scan=n_scan_sample, y_source=oracle.Y_LYOT + 300)
    t.lambda_ = oracle.ghz_to_m(GHZ)
    t.k = 2 * np.pi / t.lambda_
    row = _feed_table()
    row = row[row[:, 0] == int(GHZ)][0]
    t.th_fwhp_x, t.th_fwhp_y = np.deg2rad(row[1]), np.deg2rad(row[2])
    if threads <= 0:  # all host threads, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
        threads = len(os.sched_getaffinity(0))
    n2 = GRID_N * GRID_N
    re, im, cnt = np.zeros(n2), np.zeros(n2), np.zeros(n2)
    inv, half, kk = GRID_N / (EXTENT_CM * 10), EXTENT_CM * 5, float(t.k) / 1e3 / 2
    rows_per_slab = max(1, slab_rays // n_scan_sample)
    t_trace = t_bin = 0.0
    t_all0 = time.perf_counter()
    for r0 in range(0, n_scan_sample, rows_per_slab):
        r1 = min(n_scan_sample, r0 + rows_per_slab)
        t0 = time.perf_counter()
        out = oracle.rx_to_lyot([0, 0, 0], t, ray_begin=r0 * n_scan_sample,
                                ray_end=r1 * n_scan_sample, n_threads=threads)
        t1 = time.perf_counter()
        sel = out[4] != 0
        x, z, a, p = out[0][sel], out[2][sel], out[4][sel], kk * out[3][sel]
        ix = np.floor((x + half) * inv).astype(np.int64)
        iz = np.floor((z + half) * inv).astype(np.int64)
        ok = (ix >= 0) & (ix < GRID_N) & (iz >= 0) & (iz < GRID_N)
        cell = iz[ok] * GRID_N + ix[ok]
        re += np.bincount(cell, weights=a[ok] * np.cos(p[ok]), minlength=n2)
        im += np.bincount(cell, weights=a[ok] * np.sin(p[ok]), minlength=n2)
        cnt += np.bincount(cell, minlength=n2)
        t_trace += t1 - t0
        t_bin += time.perf_counter() - t1
    t0 = time.perf_counter()
    with np.errstate(invalid="ignore", divide="ignore"):
        b = np.where(cnt > 0, (re + 1j * im) / cnt, 0).reshape(GRID_N, GRID_N)
    B = np.fft.fftshift(np.fft.fft2(np.fft.fftshift(b)))
    t_fixed = time.perf_counter() - t0
    assert np.isfinite(B).all()
    return {"rays": n_scan_sample * n_scan_sample, "cores": threads, "t_trace": t_trace,
            "t_bin": t_bin, "t_fixed": t_fixed, "t_total": time.perf_counter() - t_all0}


def amortised_rate(sample):
    """rays/s of the CPU path on the FULL bundle: per-ray cost from the sample, the fixed
    bins -> field -> 2048^2 FFT cost paid once per 1.000014e9 rays (as in the GPU step)."""
    full = float(N_SCAN) * N_SCAN
    per_ray = (sample["t_trace"] + sample["t_bin"]) / sample["rays"]
    return full / (full * per_ray + sample["t_fixed"])


def _sample_text(n_sample, smp):
    return (f"n_scan={n_sample} ({smp['rays']} rays of the same 0.4 rad cone: {smp['t_trace']:.2f} s trace "
            f"+ {smp['t_bin']:.2f} s binning per sample) -> {GRID_N}^2 bins -> {GRID_N}^2 numpy FFT "
            f"({smp['t_fixed']:.2f} s, paid once per 1.000014e9 rays in `value`); oracle/ C port of "
            "ray_trace.rx_to_lyot (scipy brentq restated), OpenMP over rays")


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.
    Each step is a bounded sample of the workload; `value` = rays/s on the full bundle with the
    per-ray cost of the sample and the fixed cost amortised as in the GPU arm (the raw
    sample rate, fixed cost unamortised, is reported beside it)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    oracle.build()
    # size the samples so that the timed steps take ~150 s in all (warm-up samples are small:
    # a CPU code has nothing to warm but the page cache of the library)
    probe = cpu_reference_sample(400)
    rate = probe["rays"] / (probe["t_trace"] + probe["t_bin"])
    budget_s = 150.0 / max(1, args.steps)
    n_sample = int(min(N_SCAN, max(400, np.sqrt(rate * budget_s))))
    for _ in range(args.warmup):
        cpu_reference_sample(400)
    samples = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        samples.append(cpu_reference_sample(n_sample))
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    tot = {k: sum(x[k] for x in samples) for k in ("rays", "t_trace", "t_bin", "t_fixed")}
    tot["t_fixed"] /= len(samples)   # one FFT per bundle
    value = amortised_rate(tot)
    smp = samples[-1]
    sample = _sample_text(n_sample, smp)
    cores = smp["cores"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(N_SCAN),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": sample,
                         "raw_sample_rays_per_s": n_sample * n_sample / dt,
                         "trace_only_rays_per_s": tot["rays"] / tot["t_trace"]},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "note": ("the unmodified reference (pure Python, scipy brentq per ray) runs at ~600-800 "
                 "rays/s on one core (BASELINE.md); this port is ~3 orders of magnitude faster"),
    }
    print(json.dumps(line), flush=True)


def other_configs(torch, ot_geo, opt_analyze, ray_trace, table):
    """Device-timed one-liners for BASELINE.json configs[0..3] (parity for each is in tests/)."""
    from sosat_optics_b200 import geometry
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            a, b = ev(), ev()
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    out = {}
    # cfg 0: sat_nearfield notebook default through the drop-in call (host arrays in and out)
    t = _tele_geo(ot_geo, opt_analyze, table, 60)
    ray_trace.getNearField(t, [0, 0, 0])  # module load / first-launch cost stays outside
    t0 = time.perf_counter()
    for _ in range(5):
        ray_trace.getNearField(t, [0, 0, 0])
    wall = (time.perf_counter() - t0) / 5
    out["cfg0_getNearField_n60"] = {"wall_ms": wall * 1e3, "rays_per_s": 3600 / wall,
                                    "reference": "5 s (600 rays/s, notebook stderr)"}
    # cfg 1: 1000^2 far field through a2b (host float64 arrays in, complex128 out)
    amp = np.random.default_rng(1).uniform(size=(1000, 1000))
    opt_analyze.a2b(amp, amp)
    t0 = time.perf_counter()
    opt_analyze.a2b(amp, amp)
    out["cfg1_a2b_1000_host"] = {"wall_ms": (time.perf_counter() - t0) * 1e3,
                                 "reference": "81 ms (BASELINE.md, numpy 1 core)"}
    # cfg 1, the notebook's own route (sat_farfield.ipynb cells 1-6): getNearField -> sim2d (cubic
    # griddata restated on the device) -> sim_data (bicubic spline) -> zero_pad -> a2b, n_scan = 100
    t = _tele_geo(ot_geo, opt_analyze, table, 100)
    def notebook_route():
        sb = ray_trace.getNearField(t, [0, 0, 0])
        t1 = time.perf_counter()
        sb2 = ray_trace.sim2d(sb)
        t2 = time.perf_counter()
        dd = ray_trace.sim_data(sb2, 80, 2)
        t3 = time.perf_counter()
        _, _, bd = opt_analyze.zero_pad(dd.x_data, dd.y_data, dd.a_data, 100)
        xd, yd, pd = opt_analyze.zero_pad(dd.x_data, dd.y_data, dd.p_data, 100)
        bf, _ = opt_analyze.a2b(bd, np.rad2deg(pd))
        xa, ya = opt_analyze.coords_spat_to_ang(xd / 1e2, yd / 1e2, GHZ)
        return t2 - t1, t3 - t2, ray_trace.ff_fwhm_est(bf, xa, ya)
    notebook_route()
    t0 = time.perf_counter()
    w_sim2d, w_sim_data, fwhm = notebook_route()
    out["cfg1_notebook_route_n100"] = {
        "wall_ms": (time.perf_counter() - t0) * 1e3, "sim2d_ms": w_sim2d * 1e3, "sim_data_ms": w_sim_data * 1e3,
        "fwhm_arcmin": float(fwhm),
        "reference": "getNearField 15 s + sim2d 0.46 s + sim_data 0.05 s + a2b (BASELINE.md, 1 core)"}
    # cfg 2: 64-frequency sweep 77-106 GHz at n_scan=150, one fused launch
    t = _tele_geo(ot_geo, opt_analyze, table, 150)
    freqs = geometry.freq_specs_from_table(np.linspace(77, 106, 64), table)
    ms = timed(lambda: ray_trace.trace_bin_device(t, [[0, 0, 0]], freqs, 256, 420.0))
    out["cfg2_sweep_64freq_n150"] = {"ms": ms, "ray_freq_pairs_per_s": 22500 * 64 / (ms * 1e-3),
                                     "reference": "64 x 32 s (re-traces every frequency)"}
    # cfg 3: 8 x 8 focal-plane receiver sweep (|rx| <= 150 mm), n_scan = 1000, one launch
    g = np.linspace(-150, 150, 8)
    rxs = [[x, 0.0, z] for x in g for z in g if np.hypot(x, z) <= 150.0]
    t = _tele_geo(ot_geo, opt_analyze, table, 1000)
    spec = [(t.k, t.th_fwhp_x, t.th_fwhp_y)]
    ms = timed(lambda: ray_trace.trace_bin_device(t, rxs, spec, 512, 420.0), reps=3)
    out["cfg3_receiver_sweep_n1000"] = {"ms": ms, "receivers": len(rxs),
                                        "rays_per_s": len(rxs) * 1e6 / (ms * 1e-3)}
    # cfg 3 near + far field in one call: fused trace+bin, batched 1024^2 far fields (512^2 near
    # field zero-padded), on-device FWHM search; host wall time incl. the scalar read-backs
    ray_trace.beam_sweep(t, rxs[:16], spec, grid_n=512, extent_cm=42.0, pad=256)
    t0 = time.perf_counter()
    res = ray_trace.beam_sweep(t, rxs[:16], spec, grid_n=512, extent_cm=42.0, pad=256)
    wall = time.perf_counter() - t0
    out["cfg3_beam_sweep_16rx_near_plus_far_1024"] = {
        "wall_ms": wall * 1e3, "receivers": 16, "rays": 16e6,
        "fwhm_arcmin_boresight_like": float(np.median(res["fwhm_arcmin"]))}
    # configs[2] x configs[3]: 32 receivers x 64 frequencies (2 048 near + far fields, 256^2) in
    # one call -- one fused trace+bin launch, one normalisation launch, one batched transform,
    # one batched FWHM search, O(1) host synchronisations
    t = _tele_geo(ot_geo, opt_analyze, table, 150)
    freqs = geometry.freq_specs_from_table(np.linspace(77, 106, 64), table)
    ray_trace.beam_sweep(t, rxs[:32], freqs, grid_n=256, extent_cm=42.0)
    t0 = time.perf_counter()
    res = ray_trace.beam_sweep(t, rxs[:32], freqs, grid_n=256, extent_cm=42.0)
    wall = time.perf_counter() - t0
    out["cfg2x3_beam_sweep_32rx_64freq_256"] = {
        "wall_ms": wall * 1e3, "fields": 32 * 64, "fields_per_s": 32 * 64 / wall,
        "fwhm_arcmin_median": float(np.median(res["fwhm_arcmin"]))}
    opt_analyze.release_workspaces()
    # the drop-in per-ray route end to end (host arrays out): rx_to_lyot's out[17, n] table
    # (136 B/ray of D2H) and getNearField, at the notebook size and at 1.7e7 rays
    for n_scan in (150, 4096):
        t = _tele_geo(ot_geo, opt_analyze, table, n_scan)
        ray_trace.rx_to_lyot([0, 0, 0], t)
        if n_scan <= 1000:
            ray_trace.getNearField(t, [0, 0, 0])   # first use of torch's compaction at this size
        t0 = time.perf_counter()
        o = ray_trace.rx_to_lyot([0, 0, 0], t)
        w1 = time.perf_counter() - t0
        nbytes = o.nbytes
        del o
        t0 = time.perf_counter()
        sb = ray_trace.getNearField(t, [0, 0, 0])
        w2 = time.perf_counter() - t0
        out[f"dropin_e2e_n{n_scan}"] = {
            "rays": n_scan * n_scan,
            "rx_to_lyot": {"wall_ms": w1 * 1e3, "rays_per_s": n_scan * n_scan / w1,
                           "d2h_bytes": nbytes},
            "getNearField": {"wall_ms": w2 * 1e3, "rays_per_s": n_scan * n_scan / w2,
                             "d2h_bytes": int(4 * 8 * sb.a_sim.size + 24)},
            "note": "host-array API, D2H inside the timed call; bounded by 136 B/ray of table "
                    "read-back at large n_scan (pageable copy above 1 GiB)"}
        del sb
    return out


def cfg3_sweep(torch, dist, sdist, ot_geo, opt_analyze, ray_trace, table, world, rank):
    """BASELINE.json configs[3]: focal-plane receiver sweep, near + far field, receivers dealt
    round-robin to the ranks (no data-path collective; scalars gathered).  All ranks call it."""
    g = np.linspace(-150, 150, 20)
    rxs = np.array([[x, 0.0, z] for x in g for z in g if np.hypot(x, z) <= 150.0])
    t = _tele_geo(ot_geo, opt_analyze, table, 1000)
    spec = [(t.k, t.th_fwhp_x, t.th_fwhp_y)]
    kw = dict(grid_n=512, extent_cm=42.0, pad=256)

    def run():
        return sdist.beam_sweep_distributed(t, rxs, spec, **kw)

    run()
    torch.cuda.synchronize()
    times = []
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = run()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    tt = torch.tensor([min(times)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    wall = tt.item()
    opt_analyze.release_workspaces()
    if rank != 0:
        return None
    fw = res["fwhm_arcmin"][:, 0]
    return {"receivers": len(rxs), "n_scan": 1000, "rays": len(rxs) * 1e6, "near_grid": 512,
            "far_grid": 1024, "n_gpus": world, "wall_ms": wall * 1e3,
            "receivers_per_s": len(rxs) / wall, "rays_per_s": len(rxs) * 1e6 / wall,
            "fwhm_arcmin_min_median_max": [float(fw.min()), float(np.median(fw)), float(fw.max())],
            "checksum_n_valid": float(res["n_valid"].sum()),
            "sharding": "receivers round-robin over ranks, no data-path collective, scalars "
                        "gathered (host wall time incl. the gather, max over ranks, best of 3)"}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from sosat_optics_b200 import _lib, dist as sdist, opt_analyze, ot_geo, ray_trace

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: sosat_optics_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.require_gpu()
    peaks, peaks_kind = _peaks()

    table = _feed_table()
    n_scan = args.n_scan
    t = _tele_geo(ot_geo, opt_analyze, table, n_scan)
    spec = [(t.k, t.th_fwhp_x, t.th_fwhp_y)]
    rx = np.zeros((1, 3))
    # N > 1: 16-row tile rows of the launch grid dealt cyclically (every rank gets the same mix of
    # central and marginal rays; contiguous blocks left the slowest rank 10 % behind at 8 GPUs)
    my_rows = sdist.partition_tile_rows(n_scan, world, rank)
    d_rows = torch.from_numpy(my_rows).to("cuda") if world > 1 else None
    total_rays = n_scan * n_scan
    my_rays = int(sum(min(16 * int(t) + 16, n_scan) - 16 * int(t) for t in my_rows)) * n_scan

    # N > 1: the bin planes live in a buffer that every rank of the node maps (CUDA IPC), and one
    # kernel per rank reduces its slab of cells from all peers over NVLink, normalises it and
    # stores it into rank 0's field (csrc/peer.cu); --reduce nccl selects the all-reduce path
    use_peer = world > 1 and args.reduce == "peer"
    peer_note = None
    if use_peer:
        # all ranks must take the same path: agree on whether the peer mapping worked everywhere
        peer, err = None, ""
        try:
            peer = sdist.PeerPlanes(1, 1, GRID_N)
        except Exception as exc:  # no P2P between these devices, IPC disabled in the container ...
            err = repr(exc)
        ok = torch.tensor([1.0 if peer is not None else 0.0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() < 1.0:
            if peer is not None:
                peer.close()
            use_peer = False
            peer_note = "peer mapping unavailable on some rank, NCCL all-reduce used instead: " + err
            print(f"[bench rank {rank}] {peer_note}", file=sys.stderr, flush=True)
    if use_peer:
        re, im, cnt, stats, planes_flat = peer.re, peer.im, peer.cnt, peer.stats, peer.accum
    else:
        re, im, cnt, planes_flat = sdist.alloc_planes(1, 1, GRID_N)  # one buffer: one collective
        stats = torch.zeros((1, _lib.NUM_STATS), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    host_B = torch.empty((GRID_N, GRID_N), dtype=torch.complex64).pin_memory()
    # N > 1, peer path: ONE host array shared by the ranks (POSIX shared memory, page-locked in
    # every rank); each rank ships its slab of rank 0's far field over its own PCIe link
    shared_B = sdist.SharedHostArray((1, 1, GRID_N, GRID_N), np.complex64) if use_peer else None
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def step(e2e=False, kernel_events=None):
        """One pass of the hot path.  Returns the far field (device tensor)."""
        planes_flat.zero_()
        if not use_peer:
            stats.zero_()
        if kernel_events is not None:
            kernel_events[0].record()
        ray_trace.trace_bin_device(t, rx, spec, GRID_N, EXTENT_CM * 10, tile_rows=d_rows,
                                   planes=(re, im, cnt), stats=stats)
        if kernel_events is not None:
            kernel_events[1].record()
        if use_peer:
            b = peer.reduce_field(spec, dest=[0])  # the job has ONE far field: rank 0 makes it
        else:
            sdist.reduce_bins(re, im, cnt, stats, flat=planes_flat)
            b, _ = ray_trace.bins_to_field_device(re, im, cnt, stats, spec, sync=False)
        if kernel_events is not None:
            kernel_events[2].record()
        B = None
        if use_peer:
            if rank == 0:
                B = opt_analyze.far_field(b[0, 0], device=True, out=peer.out[0, 0])
        else:
            B = opt_analyze.far_field(b[0, 0], device=True)
        if kernel_events is not None:
            kernel_events[3].record()
        if e2e:
            if use_peer:
                peer.gather_to_host(shared_B)   # collective; returns when every slab has landed
            else:
                host_B.copy_(B, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        return B

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up, then exactly K timed steps ----
    for _ in range(args.warmup):
        step()
    barrier()
    # N > 1: rank 0 also computes the far field of every step.  Give it a smaller share of the
    # rays, so that all ranks reach the reduce together and rank 0's far field of step k runs
    # while the others already trace step k + 1 (weighted deal of the tile rows, calibrated on
    # three untimed steps: shares f0 = f - ff/T, f = (1 + ff/T)/N, T = trace time of the bundle).
    weights = None
    if world > 1 and args.balance:
        cal = []
        for _ in range(3):
            ke = [ev() for _ in range(4)]
            step(kernel_events=ke)
            torch.cuda.synchronize()
            cal.append([ke[0].elapsed_time(ke[1]), ke[2].elapsed_time(ke[3]) if rank == 0 else 0.0])
        c = torch.tensor(np.median(np.array(cal), axis=0), dtype=torch.float64, device="cuda")
        dist.all_reduce(c)
        t_all, t_ff = c.tolist()
        if t_all > 0 and 0 < t_ff < t_all / world:
            f = (1.0 + t_ff / t_all) / world
            weights = [f - t_ff / t_all] + [f] * (world - 1)
            my_rows = sdist.partition_tile_rows(n_scan, world, rank, weights=weights)
            d_rows = torch.from_numpy(my_rows).to("cuda")
            my_rays = int(sum(min(16 * int(tr) + 16, n_scan) - 16 * int(tr) for tr in my_rows)) * n_scan
        for _ in range(2):
            step()
        barrier()
    sampler = ClockSampler(local)
    sampler.start()
    step_ms, trace_ms, ff_ms, red_ms = [], [], [], []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)                      # L2 flush between timed iterations (256 MiB write)
        e0, e1 = ev(), ev()
        ke = [ev() for _ in range(4)]
        e0.record()
        step(kernel_events=ke)
        e1.record()
        torch.cuda.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        trace_ms.append(ke[0].elapsed_time(ke[1]))
        ff_ms.append(ke[2].elapsed_time(ke[3]))
        red_ms.append(ke[1].elapsed_time(ke[2]))
    barrier()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.result()

    dev_ms = sum(step_ms)
    agg = torch.tensor([dev_ms, sum(trace_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
    dev_ms_max, trace_ms_max = agg.tolist()
    agg_min = torch.tensor([sum(trace_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(agg_min, op=dist.ReduceOp.MIN)
    trace_ms_min = agg_min.item()
    ms_per_step = dev_ms_max / args.steps
    value = total_rays / (ms_per_step * 1e-3)

    # ---- end to end through the public host-buffer API ----
    for _ in range(2):
        step(e2e=True)
    barrier()
    e2e_t = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        ray_trace._geom_cache.clear()     # parameters are re-marshalled and re-uploaded each step
        step(e2e=True)
        e2e_t.append(time.perf_counter() - t0)
    e2e_ms = torch.tensor([sum(e2e_t) * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = total_rays / (e2e_ms.item() * 1e-3 / args.steps)
    h2d = 520 + 24 + 1536 + 8 * 8       # geometry, rx, frequency table, zeroed statistics
    d2h = GRID_N * GRID_N * 8 + 8 * 8   # far field (complex64) + statistics read by the wrapper

    # ---- N > 1: is the sharded result RIGHT?  Rank 0 recomputes the whole bundle on its own GPU
    # and compares: the ray counts are exact integers, the reduced field must agree to 1e-4 of
    # its peak (fp32 accumulation order is the only difference).
    parity_n = None
    if world > 1:
        step(e2e=True)                            # leaves the reduced statistics / field behind
        torch.cuda.synchronize()
        if use_peer:
            b_multi = peer.field[0, 0].clone()
            st_multi = peer.stats_sum[0].clone()
        else:
            b_multi = ray_trace.bins_to_field_device(re, im, cnt, stats, spec, sync=False)[0][0, 0].clone()
            st_multi = stats[0].clone()
        dist.barrier()
        if rank == 0:
            re1, im1, cnt1, st1 = ray_trace.trace_bin_device(t, rx, spec, GRID_N, EXTENT_CM * 10)
            b1 = ray_trace.bins_to_field_device(re1, im1, cnt1, st1, spec, sync=False)[0][0, 0]
            h1, hm = st1[0].cpu().numpy(), st_multi.cpu().numpy()
            peak = b1.abs().max().item()
            err = (b_multi - b1).abs().max().item() / peak
            ints_ok = all(h1[q] == hm[q] for q in (_lib.STAT_N_VALID, _lib.STAT_N_BINNED,
                                                   _lib.STAT_N_BRACKET_FAIL))
            # N_SLOW (rays sent to the bracketed solve) is a diagnostic: whether a ray within
            # rounding of the conic-edge criterion is re-solved depends on which seeding its tile
            # used, which depends on the deal of the tile rows; the RESULT of such a ray does not
            slow_diff = abs(h1[_lib.STAT_N_SLOW] - hm[_lib.STAT_N_SLOW])
            ints_ok = ints_ok and slow_diff <= 1e-4 * max(1.0, h1[_lib.STAT_N_SLOW])
            opl_ok = abs(h1[_lib.STAT_SUM_OPL] - hm[_lib.STAT_SUM_OPL]) <= 1e-11 * abs(h1[_lib.STAT_SUM_OPL])
            cnt_ok = float(cnt1.sum(dtype=torch.float64).item()) == hm[_lib.STAT_N_BINNED]
            B1 = opt_analyze.far_field(b1, device=True)
            Bm = opt_analyze.far_field(b_multi, device=True)
            ff_err = (Bm - B1).abs().max().item() / B1.abs().max().item()
            host_ok = None
            if use_peer:   # the far field every rank shipped slab-wise to the shared host array
                host_ok = bool(np.array_equal(shared_B.view()[0, 0], peer.out[0, 0].cpu().numpy()))
            ok = ints_ok and opl_ok and cnt_ok and err < 1e-4 and ff_err < 1e-4 and host_ok is not False
            parity_n = {"status": "ok" if ok else "FAILED", "n_valid": float(hm[_lib.STAT_N_VALID]),
                        "n_binned": float(hm[_lib.STAT_N_BINNED]), "counts_equal_1gpu": bool(ints_ok and cnt_ok),
                        "n_slow": float(hm[_lib.STAT_N_SLOW]), "n_slow_diff_vs_1gpu": float(slow_diff),
                        "sum_opl_rel_diff": float(abs(h1[_lib.STAT_SUM_OPL] - hm[_lib.STAT_SUM_OPL]) / abs(h1[_lib.STAT_SUM_OPL])),
                        "near_field_max_err_of_peak": err, "far_field_max_err_of_peak": ff_err,
                        "shared_host_far_field_equals_device": host_ok,
                        "how": "rank 0 re-traced the whole 1.000014e9-ray bundle on one GPU after the timed region"}
            del re1, im1, cnt1, b1, B1, Bm
        dist.barrier()

    # ---- configs[3]: receiver sweep sharded over the ranks (all ranks take part) ----
    cfg3 = cfg3_sweep(torch, dist, sdist, ot_geo, opt_analyze, ray_trace, table, world, rank)

    if rank != 0:
        if world > 1:
            dist.barrier()
            if use_peer:
                shared_B.close()
                peer.close()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (trace_bin_kernel): FP64 pipe ----
    fl, ms = __import__("ctypes").c_double(), __import__("ctypes").c_double()
    _lib.check(lib.sosat_measure_fp64_peak(__import__("ctypes").byref(fl),
                                           __import__("ctypes").byref(ms)))
    fp64_peak_tf = fl.value / 1e12
    trace_ms_step = sum(trace_ms) / args.steps     # this rank's launches (its own share of the rays)
    rays_per_s_kernel = my_rays / (trace_ms_step * 1e-3)
    ncu = {}
    try:
        with open(os.path.join(ROOT, "profiles", "trace_bin_ncu.json")) as f:
            ncu = json.load(f)
    except Exception:
        pass
    # Work the kernel EXECUTES per ray, from the committed ncu capture of this kernel (DFMA = 2
    # flop, DMUL / DADD = 1): what the fp64 pipe really does.  SURVEY 8d's 1 300 flop/ray is the
    # cost model of a 4-step fp64 Newton tracer; this kernel's algorithm needs about half of it,
    # so the survey's index (reported beside it) exceeds 1 and is not a utilisation.
    dp_flop_per_ray = ncu.get("dp_flop_per_ray", 645.0)
    achieved_tf = rays_per_s_kernel * dp_flop_per_ray / 1e12
    survey_tf = rays_per_s_kernel * FLOP_PER_RAY / 1e12
    sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    # hardware figure: ncu's dfma peak_sustained is 64 lanes per SM and cycle (half rate)
    hw_fp64_tf = 2.0 * 64 * sm_count * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
    traffic = ncu.get("dram_bytes_per_launch")
    roofline = {
        "kernel": "trace_bin_kernel",
        "bound": "fp64 pipe through the issue port: a DP warp instruction holds an SMSP's dispatch for "
                 "~2 cycles (half-rate pipe), so DP instructions x 2 + other instructions ~ cycles",
        "achieved": achieved_tf, "peak": fp64_peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak_tf,
        "traffic": traffic, "share_of_step": trace_ms_step / ms_per_step,
        "algorithmic": f"{dp_flop_per_ray:.0f} executed fp64 flop/ray (2 x DFMA + DMUL + DADD per ray from the "
                       f"ncu capture) x {my_rays} rays per launch; 0 B/ray of HBM traffic (+12 B per flushed "
                       "bin: 3 planes x 2048^2 x 4 B = 50 MB per launch)",
        "peak_source": "FP64 FMA microkernel measured in this run (sosat_measure_fp64_peak); "
                       "MEASURED_PEAKS.json carries no FP64 figure",
        "peak_hw_tflops": hw_fp64_tf, "frac_of_hw_peak": achieved_tf / hw_fp64_tf,
        "survey_index": {"flop_per_ray": FLOP_PER_RAY, "tflops": survey_tf, "over_measured_peak": survey_tf / fp64_peak_tf,
                         "note": "SURVEY 8d's Newton-4 cost model x rays/s: a throughput index, > 1 "
                                 "because the algorithm does less work per ray than the model"},
        "pipe_fp64_util": ncu.get("pipe_fp64_util"), "issue_util": ncu.get("issue_util"),
        "pipe_xu_util": ncu.get("pipe_xu_util"),
        "inst_per_ray": ncu.get("inst_per_ray"), "dp_inst_per_ray": ncu.get("dp_inst_per_ray"),
        "ncu_capture": ncu.get("capture"),
    }

    # ---- far-field transform time at 1024^2 and 2048^2 (metric 2) ----
    farfield = {}
    import ctypes as _ct
    for n in (1024, 2048, 4096):
        d = torch.randn(n, n, dtype=torch.complex64, device="cuda")
        o = torch.empty_like(d)
        ws = opt_analyze._workspace(lib, torch, n)
        c_args = (_ct.c_void_p(d.data_ptr()), n, _ct.c_void_p(o.data_ptr()), _ct.c_void_p(ws.data_ptr()),
                  ws.numel(), _ct.c_void_p(torch.cuda.current_stream().cuda_stream))

        def single_shot(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(20):
                a, b_ = ev(), ev()
                a.record()
                fn()
                b_.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b_))
            return min(ts), statistics.median(ts)

        # metric 2: device-resident b -> device-resident B.  Timed at the product boundary, the
        # C-ABI call sosat_dft2_gemm (three launches: transpose/split, two UMMA passes), and
        # through the Python wrapper (which adds ~15 us of host work in front of the first launch)
        best, med = single_shot(lambda: _lib.check(lib.sosat_dft2_gemm(*c_args)))
        w_best, w_med = single_shot(lambda: opt_analyze.far_field(d, device=True))
        a, b_ = ev(), ev()
        a.record()
        for _ in range(20):  # back to back: launch latency of the 3-kernel chain hidden
            lib.sosat_dft2_gemm(*c_args)
        b_.record()
        torch.cuda.synchronize()
        tf = 16.0 * n ** 3 / (best * 1e-3) / 1e12
        farfield[f"n{n}"] = {"ms_best": best, "ms_median": med, "timed": "C-ABI call sosat_dft2_gemm, single shot",
                             "ms_python_wrapper_best": w_best, "ms_python_wrapper_median": w_med,
                             "ms_back_to_back": a.elapsed_time(b_) / 20, "algorithmic_tflops": tf,
                             "frac_of_bf16_peak": tf / peaks["bf16_tflops"],
                             "executed_tflops": 3 * tf}
        del d, o
    # sweep form: 16 fields of 1024^2 through one pair of UMMA launches (far_field_batch)
    d = torch.randn(16, 1024, 1024, dtype=torch.complex64, device="cuda")
    for _ in range(3):
        opt_analyze.far_field_batch(d, device=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b_ = ev(), ev()
        a.record()
        opt_analyze.far_field_batch(d, device=True)
        b_.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b_))
    tf = 16.0 * 1024 ** 3 * 16 / (min(ts) * 1e-3) / 1e12
    farfield["n1024_batch16"] = {"ms_best": min(ts), "ms_per_transform": min(ts) / 16,
                                 "algorithmic_tflops": tf,
                                 "frac_of_bf16_peak": tf / peaks["bf16_tflops"]}
    del d
    opt_analyze.release_workspaces()
    farfield["roofline"] = {
        "kernel": "prep_b_c64_kernel + 2 UMMA passes: gemm2_bf16x3_kernel (cta_group::2, 256x256 pair "
                  "tiles) at 2048^2, gemm_bf16x3_kernel (2x2-cluster TMA multicast) at 1024^2; PDL",
        "bound": "tensor",
        "achieved": farfield["n2048"]["algorithmic_tflops"], "peak": peaks["bf16_tflops"],
        "unit": "TFLOP/s", "frac": farfield["n2048"]["frac_of_bf16_peak"],
        "peak_source": f"{peaks_kind} bf16 burst (MEASURED_PEAKS.json)",
        "note": "16 N^3 algorithmic flop; the bf16x3 split executes 3x that on the tensor pipe",
    }

    # ---- K3: direct-sum far field on irregular samples (SFU bound: 2 MUFU per point-angle pair) ----
    if world == 1:
        import ctypes
        n_pts, n_ang = 100_000, 256 * 256
        gen = torch.Generator(device="cuda").manual_seed(3)
        pts = torch.rand((n_pts, 4), device="cuda", generator=gen) - 0.5
        th = (torch.rand((n_ang, 2), device="cuda", generator=gen) - 0.5) * 0.05
        outB = torch.empty((n_ang, 2), dtype=torch.float32, device="cuda")
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

        def nudft():
            _lib.check(lib.sosat_nudft_direct(ctypes.c_void_p(pts.data_ptr()), n_pts,
                                              ctypes.c_void_p(th.data_ptr()), n_ang, 1.0 / 3.33e-3,
                                              ctypes.c_void_p(outB.data_ptr()), st))
        for _ in range(3):
            nudft()
        torch.cuda.synchronize()
        k3_sampler = ClockSampler(local)   # this kernel keeps the SFU pipe > 90 % busy: watch the clocks
        k3_sampler.start()
        ts = []
        for _ in range(12):
            a, b_ = ev(), ev()
            a.record()
            nudft()
            b_.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b_))
        k3_clocks = k3_sampler.result()
        pairs = n_pts * n_ang / (min(ts) * 1e-3)
        sm_count = torch.cuda.get_device_properties(local).multi_processor_count
        mufu_peak = 16.0 * sm_count * (clocks.get("sm_max_mhz") or 1965.0) * 1e6  # MUFU lanes/clk/SM
        farfield["nudft_direct"] = {
            "points": n_pts, "angles": n_ang, "ms_best": min(ts), "ms_median": statistics.median(ts),
            "clocks": k3_clocks, "pairs_per_s": pairs,
            "bound": "sfu", "achieved_mufu_per_s": 2 * pairs, "peak_mufu_per_s": mufu_peak,
            "frac": 2 * pairs / mufu_peak,
            "note": "peak quoted at the maximum SM clock; this kernel runs into sw_power_cap on B200 (see clocks: "
                    "the SM clock drops to ~1.4-1.6 GHz under it), so ms_best varies 3.1-4.4 ms between boxes. "
                    "2 MUFU (sin, cos) + 7 FP32 per pair, points split over gridDim.y; 16 MUFU lanes/clk/SM (measured: "
                    "8 cycles per warp instruction and SMSP, scripts/dp_latency.cu)"}

    # ---- the other BASELINE.json configs, one line each (N = 1 only; not the headline) ----
    configs = None
    if world == 1:
        configs = other_configs(torch, ot_geo, opt_analyze, ray_trace, table)
        for n in (1024, 2048):   # reference far field on the host: opt_analyze.a2b = shifted fft2
            fld = (np.random.default_rng(2).standard_normal((n, n)) + 0j)
            t0 = time.perf_counter()
            np.fft.fftshift(np.fft.fft2(np.fft.fftshift(fld)))
            farfield[f"n{n}"]["cpu_numpy_fft2_ms"] = (time.perf_counter() - t0) * 1e3

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        probe = cpu_reference_sample(1000)             # sizing probe
        rate = probe["rays"] / (probe["t_trace"] + probe["t_bin"])
        n_sample = int(min(N_SCAN, max(1000, np.sqrt(rate * 25.0))))
        smp = cpu_reference_sample(n_sample)           # ~25 s of CPU work
        cpu = {"value": amortised_rate(smp), "unit": "rays/s", "cores": smp["cores"], "kind": "port",
               "sample": _sample_text(n_sample, smp),
               "raw_sample_rays_per_s": smp["rays"] / smp["t_total"],
               "trace_only_rays_per_s": smp["rays"] / smp["t_trace"],
               "reference_python_rays_per_s": "600-800 on 1 core (BASELINE.md; cannot use more)"}

    launches_per_step = 5  # trace_bin, bins_to_field, prep_b (transpose+split), gemm pass 1, gemm pass 2
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(n_scan),
        "run": {"sharding": f"16-row tile rows of the launch grid dealt cyclically to {world} rank(s), " + (
                       "fused reduce over peer-mapped memory: each rank sums its slab of the [re,im,count] "
                       "planes and the 8 fp64 sums from all peers, normalises it and stores it to rank 0, "
                       "which computes the far field" if use_peer else
                       "one NCCL all-reduce of [re,im,count] planes (50 MB) + 8 fp64 sums" + (
                           f" ({peer_note})" if peer_note else "")),
                "ray_shares": weights,
                "l2": "256 MiB flush write between timed steps; the kernel reads no input arrays",
                "newton_schedule": os.environ.get("SOSAT_TRACE_VARIANT", "default (fp32 Newton + Halley step, 1 fp64 Newton step, first-order slope update; long steps re-traced adaptively)")},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "farfield": farfield,
        "cpu_baseline": cpu,
        "parity_n": parity_n,
        "cfg3_receiver_sweep": cfg3,
        "configs": configs,
        "wall_s_timed_region": wall_s,
        # where a step's time goes on rank 0 (N > 1: what keeps the curve below linear)
        "step_breakdown_ms": {
            "trace_slowest_rank": trace_ms_max / args.steps,
            "trace_fastest_rank": trace_ms_min / args.steps,
            "trace_ideal": None if world == 1 else "t(1 GPU) / N: compare with BENCH N=1 ms_per_step",
            "reduce_incl_wait_for_slowest_rank": statistics.median(red_ms),
            "far_field_rank0": statistics.median(ff_ms),
            "e2e_minus_device_step (D2H of the far field + launch/marshal)":
                e2e_ms.item() / args.steps - ms_per_step},
        "far_field_ms_in_step": statistics.median(ff_ms),
        "reduce_ms_in_step": statistics.median(red_ms),  # rank 0: reduce (+ wait for the slowest rank) + bins -> field
        "peaks": peaks_kind,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if use_peer:
            shared_B.close()
            peer.close()
        dist.destroy_process_group()
    if parity_n is not None and parity_n["status"] != "ok":
        raise SystemExit("N > 1 parity check FAILED: " + json.dumps(parity_n))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--n-scan", type=int, default=N_SCAN, help="reduce only for debugging")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-balance", dest="balance", action="store_false",
                    help="N > 1: equal shares of the rays for all ranks (default: rank 0, which also computes "
                         "the far field, gets a smaller share, calibrated on untimed steps)")
    ap.add_argument("--reduce", choices=["peer", "nccl"], default="peer",
                    help="N > 1: fused peer-memory reduce (default) or NCCL all-reduce of the bin planes")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "ours" and world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.gpus > 1 and world == 1:
            raise SystemExit("for --gpus N > 1 launch with python -m torch.distributed.run "
                             "--nproc-per-node N (one rank per GPU)")
        run_ours(args)


if __name__ == "__main__":
    main()
