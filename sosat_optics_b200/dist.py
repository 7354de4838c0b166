"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for plumbing.

The path shards naturally:

* **rays** of one bundle: contiguous blocks of launch rows (iphi) per rank, each rank bins into
  its own full-size private grid, then ONE all-reduce(sum) of the ``[re, im, count]`` planes
  plus the 8 fp64 statistics over NVLink (NCCL) -- the only collective of the data path;
* **frequencies / receiver positions**: independent problems, dealt round-robin to ranks with
  no data-path collective (results are gathered).

Nothing here launches processes: use ``torchrun`` / ``python -m torch.distributed.run``.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

ROW_ALIGN = 16  # the tracer's CTA tile edge (csrc/trace.cu kTile): keep tiles whole


def partition_rows(n_scan: int, world: int, rank: int, align: int = ROW_ALIGN) -> Tuple[int, int]:
    """Contiguous block of launch rows [begin, end) for ``rank``; blocks are multiples of
    ``align`` rows (except the last), cover [0, n_scan) exactly and differ by <= ``align``."""
    if not (0 <= rank < world):
        raise ValueError("rank outside [0, world)")
    n_units = (n_scan + align - 1) // align
    base, extra = divmod(n_units, world)
    u0 = rank * base + min(rank, extra)
    u1 = u0 + base + (1 if rank < extra else 0)
    return min(u0 * align, n_scan), min(u1 * align, n_scan)


def partition_items(n_items: int, world: int, rank: int) -> List[int]:
    """Round-robin deal of independent problems (frequencies x receivers) to ranks."""
    if not (0 <= rank < world):
        raise ValueError("rank outside [0, world)")
    return list(range(rank, n_items, world))


def _dist():
    import torch.distributed as dist
    return dist


def world_info(group=None) -> Tuple[int, int]:
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def alloc_planes(n_rx: int, n_freq: int, grid_n: int, device="cuda"):
    """Accumulator planes ``(re, im, cnt)`` as views of ONE flat float32 buffer (returned last),
    so that :func:`reduce_bins` can all-reduce them in place with a single collective."""
    import torch
    n_ri = n_rx * n_freq * grid_n * grid_n
    n_c = n_rx * grid_n * grid_n
    flat = torch.zeros(2 * n_ri + n_c, dtype=torch.float32, device=device)
    re = flat[:n_ri].view(n_rx, n_freq, grid_n, grid_n)
    im = flat[n_ri:2 * n_ri].view(n_rx, n_freq, grid_n, grid_n)
    cnt = flat[2 * n_ri:].view(n_rx, grid_n, grid_n)
    return re, im, cnt, flat


def reduce_bins(re, im, cnt, stats, group=None, flat=None):
    """All-reduce(sum) of the near-field accumulators in place: one collective over the three
    float32 planes (50 MB at 2048^2) and a second, tiny one over the fp64 statistics.  Pass the
    ``flat`` buffer of :func:`alloc_planes` to reduce the planes where they lie; otherwise they
    are packed into a temporary and copied back."""
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return re, im, cnt, stats
    import torch
    if flat is not None:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    else:
        packed = torch.cat([re.reshape(-1), im.reshape(-1), cnt.reshape(-1)])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        n_re, n_im = re.numel(), im.numel()
        re.copy_(packed[:n_re].view_as(re))
        im.copy_(packed[n_re:n_re + n_im].view_as(im))
        cnt.copy_(packed[n_re + n_im:].view_as(cnt))
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return re, im, cnt, stats


def near_field_binned_distributed(tele_geo, rx, grid_n: int, extent_cm: float,
                                  freqs: Optional[Sequence] = None, group=None,
                                  trace_fn: Optional[Callable] = None,
                                  field_fn: Optional[Callable] = None):
    """Ray-sharded fused trace + binning: every rank traces its block of launch rows, the
    bin planes are all-reduced, and every rank ends with the full field.

    ``trace_fn`` / ``field_fn`` default to the CUDA entry points
    (:func:`ray_trace.trace_bin_device`, :func:`ray_trace.bins_to_field_device`); they are
    injectable so that the sharding and the reduce can be exercised on CPU with ``gloo``.
    Returns ``(b, cnt, stats)``: complex field ``[n_rx, n_freq, N, N]``, counts, host stats."""
    from . import ray_trace
    rank, world = world_info(group)
    trace_fn = trace_fn or ray_trace.trace_bin_device
    field_fn = field_fn or ray_trace.bins_to_field_device
    specs = list(freqs) if freqs is not None else [
        (tele_geo.k, tele_geo.th_fwhp_x, tele_geo.th_fwhp_y)]
    r0, r1 = partition_rows(int(tele_geo.n_scan), world, rank)
    rx_arr = np.asarray(rx, dtype=np.float64).reshape(-1, 3)
    re, im, cnt, stats = trace_fn(tele_geo, rx_arr, specs, grid_n, extent_cm * 10.0,
                                  row_begin=r0, row_end=r1)
    re, im, cnt, stats = reduce_bins(re, im, cnt, stats, group)
    b, h_stats = field_fn(re, im, cnt, stats, specs)
    return b, cnt, h_stats


def sweep_distributed(problems: Sequence, solve: Callable, group=None) -> list:
    """Frequency / receiver sweeps: ``solve(problem)`` runs on the rank that owns the problem
    (round-robin); the (small, picklable) results are gathered to every rank in problem order.
    No data-path collective."""
    dist = _dist()
    rank, world = world_info(group)
    mine = {i: solve(problems[i]) for i in partition_items(len(problems), world, rank)}
    if world == 1:
        return [mine[i] for i in range(len(problems))]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    merged = {}
    for part in gathered:
        merged.update(part)
    return [merged[i] for i in range(len(problems))]
