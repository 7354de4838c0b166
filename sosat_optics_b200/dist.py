"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for plumbing.

The path shards naturally:

* **rays** of one bundle: 16-row tile rows of the launch grid dealt cyclically to the ranks (or
  contiguous row blocks, :func:`partition_rows`), each rank bins into its own full-size private
  grid, then ONE reduce of the ``[re, im, count]`` planes plus the 8
  fp64 statistics over NVLink -- the only exchange of the data path.  Two implementations:
  :func:`reduce_bins` (NCCL all-reduce, then the normalisation kernel) and
  :class:`PeerPlanes` (one kernel per rank over peer-mapped memory that reduces its slab of
  cells, normalises it and stores it to every rank: ``csrc/peer.cu``);
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


def partition_tile_rows(n_scan: int, world: int, rank: int, align: int = ROW_ALIGN,
                        weights: Optional[Sequence[float]] = None) -> np.ndarray:
    """Tile rows (``align`` launch rows each) of ``rank`` under a CYCLIC deal: rows rank,
    rank + world, ...  Every rank then traces the same mix of central rays (all valid, all
    binned, predicted seeds) and marginal rays (clipped, or outside lens 1b's conic domain: no
    binning but fp32-seeded tiles), which contiguous blocks do not give: at 8 GPUs the slowest
    block took 5.15 ms against 4.61 ms for the fastest.

    ``weights`` (one positive number per rank, identical on all ranks): shares of the rays, dealt
    by weighted round-robin -- e.g. a smaller share for the rank that also computes the far field
    of every step, so that all ranks reach the reduce at the same time."""
    if not (0 <= rank < world):
        raise ValueError("rank outside [0, world)")
    n_rows = (n_scan + align - 1) // align
    if weights is None:
        return np.arange(rank, n_rows, world, dtype=np.int32)
    w = np.asarray(weights, dtype=np.float64)
    if w.shape != (world,) or not np.all(w > 0) or not np.all(np.isfinite(w)):
        raise ValueError("weights: one positive finite number per rank")
    # row t goes to the rank whose next row comes first on its own clock (stride 1 / weight);
    # equal weights give the cyclic deal
    nxt = 1.0 / w - 1.0 / w.max()     # phase: the heaviest rank starts
    mine = []
    for t in range(n_rows):
        r = int(np.argmin(nxt))
        if r == rank:
            mine.append(t)
        nxt[r] += 1.0 / w[r]
    return np.asarray(mine, dtype=np.int32)


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


class _RawDeviceBuffer:
    """``__cuda_array_interface__`` holder for a buffer owned by libsosat_b200 (zero-copy view)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1",
                                         "data": (ptr, False), "version": 2}


def peer_layout(n_rx: int, n_freq: int, grid_n: int) -> dict:
    """Byte offsets of ``[re | im | cnt | stats | field]`` inside a :class:`PeerPlanes` buffer
    (every region 256-byte aligned; identical on all ranks)."""
    n_ri = n_rx * n_freq * grid_n * grid_n
    n_c = n_rx * grid_n * grid_n
    off, out = 0, {}
    for name, nbytes in (("re", 4 * n_ri), ("im", 4 * n_ri), ("cnt", 4 * n_c),
                         ("stats", 8 * 8 * n_rx), ("stats_sum", 8 * 8 * n_rx), ("field", 8 * n_ri),
                         ("out", 8 * n_ri), ("flags", 64 * 8)):
        out[name] = off
        off += (nbytes + 255) // 256 * 256
    out["bytes"] = off
    out["accum_bytes"] = out["stats_sum"]  # [re | im | cnt | stats]: what a step zeroes
    return out


def slab_cells(n_cells: int, world: int, rank: int, align: int = 4) -> Tuple[int, int]:
    """Cells [begin, end) of one field that ``rank`` reduces (begin a multiple of ``align``)."""
    return partition_rows(n_cells, world, rank, align)


class PeerPlanes:
    """Accumulator planes of every rank of the node, mapped into this rank's address space.

    ``re, im, cnt, stats`` are torch views of the local buffer (pass them to
    ``ray_trace.trace_bin_device``); :meth:`reduce_field` is the fused reduce + normalisation
    + gather.  One process per GPU on ONE node (CUDA IPC over NVLink / PCIe peer access)."""

    def __init__(self, n_rx: int, n_freq: int, grid_n: int, group=None):
        import ctypes
        import torch
        from . import _lib
        self.lib = _lib.require_gpu()
        self.group = group
        self.rank, self.world = world_info(group)
        if self.world > 16:
            raise ValueError("PeerPlanes supports at most 16 ranks (one node)")
        if n_rx * n_freq > 1 and (grid_n * grid_n) % 4:
            raise ValueError("PeerPlanes with more than one field needs grid_n**2 to be a multiple "
                             f"of 4 (16-byte aligned planes for the P2P loads); got grid_n={grid_n}")
        self.shape = (n_rx, n_freq, grid_n)
        self.layout = L = peer_layout(n_rx, n_freq, grid_n)
        # Every step that can fail locally (allocation, export, mapping a peer) is followed by
        # an exchange of the outcome, so that all ranks raise together instead of one rank
        # raising while the others wait in a collective.
        self.base, self._opened, self.peer_base = 0, [], []
        err, handle = None, (ctypes.c_ubyte * _lib.PEER_HANDLE_BYTES)()
        try:
            base = ctypes.c_void_p()
            _lib.check(self.lib.sosat_peer_alloc(L["bytes"], ctypes.byref(base)))
            self.base = base.value
            _lib.check(self.lib.sosat_peer_export(ctypes.c_void_p(self.base), handle))
        except Exception as exc:  # noqa: BLE001 - reported to every rank below
            err = repr(exc)
        handles = self._exchange(None if err else bytes(handle))
        self._raise_together(err, [h is None for h in handles], "allocating / exporting")
        try:
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.peer_base.append(self.base)
                    continue
                ptr = ctypes.c_void_p()
                buf = (ctypes.c_ubyte * _lib.PEER_HANDLE_BYTES).from_buffer_copy(h)
                _lib.check(self.lib.sosat_peer_open(buf, ctypes.byref(ptr)))
                self.peer_base.append(ptr.value)
                self._opened.append(ptr.value)
        except Exception as exc:  # noqa: BLE001
            err = repr(exc)
        self._raise_together(err, [f is not None for f in self._exchange(err)], "mapping a peer of")
        self._raw = _RawDeviceBuffer(self.base, L["bytes"])
        self.flat = torch.as_tensor(self._raw, device="cuda")
        n_ri, n_c = n_rx * n_freq * grid_n * grid_n, n_rx * grid_n * grid_n

        def view(name, nbytes, dtype, shape):
            return self.flat[L[name]:L[name] + nbytes].view(dtype).view(shape)

        self.re = view("re", 4 * n_ri, torch.float32, (n_rx, n_freq, grid_n, grid_n))
        self.im = view("im", 4 * n_ri, torch.float32, (n_rx, n_freq, grid_n, grid_n))
        self.cnt = view("cnt", 4 * n_c, torch.float32, (n_rx, grid_n, grid_n))
        self.stats = view("stats", 64 * n_rx, torch.float64, (n_rx, 8))
        self.stats_sum = view("stats_sum", 64 * n_rx, torch.float64, (n_rx, 8))
        self.field = torch.view_as_complex(
            view("field", 8 * n_ri, torch.float32, (n_rx, n_freq, grid_n, grid_n, 2)))
        # far fields of the reduced near fields, written where every rank of the node can read them
        # (gather_to_host: each rank ships a slab to the host over its own PCIe link)
        self.out = torch.view_as_complex(
            view("out", 8 * n_ri, torch.float32, (n_rx, n_freq, grid_n, grid_n, 2)))
        self.accum = self.flat[:L["accum_bytes"]]
        self._flag = torch.zeros(1, dtype=torch.float32, device="cuda")
        self.flags = view("flags", 64 * 8, torch.int64, (64,))
        self._epoch = 0
        # barriers: flags in the peer-mapped buffers (default) or a one-element NCCL all-reduce
        import os
        self._flag_barrier = os.environ.get("SOSAT_PEER_BARRIER", "flags") != "nccl"
        self._c_peers = (ctypes.c_void_p * self.world)(*self.peer_base)

    def _exchange(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        _dist().all_gather_object(out, obj, group=self.group)
        return out

    def _raise_together(self, err, failed, what):
        if any(failed):
            self.close()
            bad = [r for r, f in enumerate(failed) if f]
            raise RuntimeError(f"PeerPlanes: {what} the accumulator buffer failed on rank(s) {bad}"
                               + (f": {err}" if err else ""))

    def zero(self):
        """Zero the accumulators (planes and statistics) for the next step."""
        self.accum.zero_()

    def barrier(self):
        """Stream-ordered barrier over the ranks, no host sync: every rank writes an epoch into
        a flag slot of every peer's buffer and waits for all of its own slots
        (``sosat_peer_barrier``); ``SOSAT_PEER_BARRIER=nccl`` selects the one-element NCCL
        all-reduce it replaced."""
        if self.world == 1:
            return
        if not self._flag_barrier:
            _dist().all_reduce(self._flag, group=self.group)
            return
        import ctypes
        import torch
        from . import _lib
        self._epoch += 1
        _lib.check(self.lib.sosat_peer_barrier(
            self._c_peers, self.world, self.rank, self.layout["flags"], self._epoch, 10.0,
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def check(self):
        """Raise if a flag barrier of this rank ever timed out (host sync; call off the hot path)."""
        if self.world > 1 and self.base and int(self.flags[63].item()) != 0:
            raise RuntimeError(f"PeerPlanes: rank {self.rank} timed out in the peer barrier of epoch "
                               f"{int(self.flags[63].item())}: a peer never arrived")

    def reduce_field(self, freq_specs, opl_ref: float = 0.0, dest: Optional[Sequence[int]] = None):
        """Sum every rank's planes and statistics, normalise to the complex field and store it
        in ``self.field`` of the ranks in ``dest`` (default: all).  Collective: every rank calls
        it after its trace.  Returns ``self.field`` (valid on the destination ranks)."""
        import ctypes
        import torch
        from . import _lib, geometry
        n_rx, n_freq, n = self.shape
        L = self.layout
        dest = list(range(self.world)) if dest is None else list(dest)
        c_out = (ctypes.c_void_p * len(dest))(*[self.peer_base[d] for d in dest])
        b0, b1 = slab_cells(n * n, self.world, self.rank)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        self.barrier()  # every rank's trace has finished
        ks = (ctypes.c_double * n_freq)(*[geometry.scalar(k) / 1e3 / 2 for k, _, _ in freq_specs])
        # one launch for the whole stack of (receiver, frequency) fields (grid y = field)
        _lib.check(self.lib.sosat_bins_reduce_fields(
            self._c_peers, self.world, L["re"], L["im"], L["cnt"], L["stats"], n * n, n_rx, n_freq,
            b0, b1, ks, float(opl_ref), c_out, len(dest), L["field"],
            ctypes.c_void_p(self.stats_sum.data_ptr()), stream))
        self.barrier()  # every slab has landed; the planes may be overwritten
        return self.field

    def gather_to_host(self, host: "SharedHostArray", root: int = 0):
        """Ship ``self.out`` of rank ``root`` (e.g. the far field it computed) to the host array
        shared by the ranks of the node: every rank reads its slab of cells from the root's
        buffer over NVLink (peer-mapped memory) and copies it to the page-locked shared segment
        over ITS OWN PCIe link -- W concurrent 1/W-size device->host copies instead of one (at 8
        GPUs: 8 x 4 MB instead of 33.5 MB through one link).  Collective; returns when every
        slab has landed (``host.array`` is then complete on every rank)."""
        import torch
        n_rx, n_freq, n = self.shape
        total = n_rx * n_freq * n * n
        if host.array.size != total or host.array.dtype != np.complex64:
            raise ValueError("host array must hold n_rx * n_freq * grid_n**2 complex64 values")
        b0, b1 = slab_cells(total, self.world, self.rank)
        self.barrier()  # the root's far field is complete (stream order + one-element all-reduce)
        if b1 > b0:
            if getattr(self, "_peer_out", None) is None:
                L = self.layout
                raw = _RawDeviceBuffer(self.peer_base[root] + L["out"], 8 * total)
                self._peer_out_raw = raw
                self._peer_out = torch.as_tensor(raw, device="cuda").view(torch.complex64)
                self._stage = torch.empty(b1 - b0, dtype=torch.complex64, device="cuda")
            self._stage.copy_(self._peer_out[b0:b1])                     # NVLink (local on the root)
            host.tensor[b0:b1].copy_(self._stage, non_blocking=True)     # this rank's PCIe link
        # stream-ordered barrier behind the copy: when this rank's all-reduce has completed, every
        # rank's contribution -- enqueued after its own device->host copy -- has run
        self.barrier()
        torch.cuda.current_stream().synchronize()

    def close(self):
        """Collective: every rank of the group calls it.  The local buffer is freed only after
        EVERY rank has unmapped it -- freeing exported memory that a peer still has open through
        ``cudaIpcOpenMemHandle`` is undefined behaviour -- so the order is: drain this device,
        close the mappings of the peers' buffers, exchange "closed" over the group (host-level
        ``all_gather_object``, as in ``__init__``), then free."""
        import ctypes
        import torch
        from . import _lib
        if getattr(self, "_closed", False):
            return
        self._closed = True
        timed_out = None
        if torch.cuda.is_available():
            torch.cuda.synchronize()  # no reduce kernel of this rank still reads a peer
            try:
                self.check()
            except RuntimeError as exc:
                timed_out = exc
        err = None
        for ptr in self._opened:
            try:
                _lib.check(self.lib.sosat_peer_close(ctypes.c_void_p(ptr)))
            except Exception as exc:  # noqa: BLE001 - still take part in the exchange below
                err = exc
        self._opened = []
        # every rank takes part, also one whose own allocation failed (base == 0)
        self._exchange("closed")  # every peer has dropped its mapping of this rank's buffer
        if self.base:
            for name in ("re", "im", "cnt", "stats", "stats_sum", "field", "out", "flags", "accum", "flat",
                         "_raw", "_peer_out"):
                if hasattr(self, name):
                    setattr(self, name, None)
            _lib.check(self.lib.sosat_peer_free(ctypes.c_void_p(self.base)))
            self.base = 0
        if err is not None:
            raise err
        if timed_out is not None:
            raise timed_out


def near_field_binned_distributed(tele_geo, rx, grid_n: int, extent_cm: float,
                                  freqs: Optional[Sequence] = None, group=None,
                                  trace_fn: Optional[Callable] = None,
                                  field_fn: Optional[Callable] = None,
                                  peer: Optional["PeerPlanes"] = None):
    """Ray-sharded fused trace + binning: every rank traces its block of launch rows, the
    bin planes are reduced, and every rank ends with the full field.

    ``trace_fn`` / ``field_fn`` default to the CUDA entry points
    (:func:`ray_trace.trace_bin_device`, :func:`ray_trace.bins_to_field_device`); they are
    injectable so that the sharding and the reduce can be exercised on CPU with ``gloo``.
    ``peer``: a :class:`PeerPlanes` of matching shape selects the fused peer-memory reduce
    instead of the NCCL all-reduce (the counts are then not gathered: ``cnt`` is ``None``).
    Returns ``(b, cnt, stats)``: complex field ``[n_rx, n_freq, N, N]``, counts, host stats."""
    from . import ray_trace
    rank, world = world_info(group)
    trace_fn = trace_fn or ray_trace.trace_bin_device
    field_fn = field_fn or ray_trace.bins_to_field_device
    specs = list(freqs) if freqs is not None else [
        (tele_geo.k, tele_geo.th_fwhp_x, tele_geo.th_fwhp_y)]
    rows = partition_tile_rows(int(tele_geo.n_scan), world, rank)  # cyclic deal of 16-row tiles
    rx_arr = np.asarray(rx, dtype=np.float64).reshape(-1, 3)
    if peer is not None:
        if peer.shape != (len(rx_arr), len(specs), grid_n):
            raise ValueError("PeerPlanes shape %r does not match the request" % (peer.shape,))
        peer.zero()
        trace_fn(tele_geo, rx_arr, specs, grid_n, extent_cm * 10.0, tile_rows=rows,
                 planes=(peer.re, peer.im, peer.cnt), stats=peer.stats)
        b = peer.reduce_field(specs).clone()
        return b, None, peer.stats_sum.cpu().numpy()
    re, im, cnt, stats = trace_fn(tele_geo, rx_arr, specs, grid_n, extent_cm * 10.0,
                                  tile_rows=rows)
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


def beam_sweep_distributed(tele_geo, rx_list, freqs, grid_n: int = 256,
                           extent_cm: Optional[float] = None, pad: int = 0, group=None,
                           sweep_fn: Optional[Callable] = None) -> dict:
    """BASELINE configs[3] over the GPUs of a node: the receivers of a focal-plane sweep are
    dealt round-robin to the ranks (:func:`partition_items`), every rank runs
    :func:`ray_trace.beam_sweep` on its receivers (one fused trace+bin launch, one batched
    far-field transform, one batched FWHM search) and the per-field scalars are gathered --
    no data-path collective.  Every rank returns the dict of ``beam_sweep`` for ALL receivers,
    in the order of ``rx_list``.  ``sweep_fn`` is injectable so that the dealing and the merge
    can be exercised on CPU with ``gloo``; it defaults to the CUDA ``beam_sweep``."""
    from . import ray_trace
    dist = _dist()
    rank, world = world_info(group)
    sweep_fn = sweep_fn or ray_trace.beam_sweep
    rx_arr = np.asarray(rx_list, dtype=np.float64).reshape(-1, 3)
    mine = partition_items(len(rx_arr), world, rank)
    part = sweep_fn(tele_geo, rx_arr[mine], freqs, grid_n, extent_cm, pad) if mine else None
    if world == 1:
        return part
    if sweep_fn is ray_trace.beam_sweep and dist.get_backend(group) == "nccl":
        return _gather_sweep_nccl(part, len(rx_arr), len(list(freqs)), world, rank, group)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, part), group=group)
    first = next(p for _, p in gathered if p is not None)
    merged = {k: np.zeros((len(rx_arr),) + np.asarray(v).shape[1:], dtype=np.asarray(v).dtype)
              for k, v in first.items()}
    for idx, p in gathered:
        if p is None:
            continue
        for k, v in p.items():
            merged[k][idx] = v
    return merged


# beam_sweep's result as one float64 row per receiver: [fwhm F | peak F | peak_index 2F | 3 scalars]
_SWEEP_KEYS = ("fwhm_arcmin", "peak", "peak_index", "n_valid", "n_binned", "mean_opl")


def _gather_sweep_nccl(part, n_rx: int, n_freq: int, world: int, rank: int, group) -> dict:
    """Gather the per-receiver scalars of a sharded sweep with ONE NCCL all-gather of a packed
    float64 tensor (an ``all_gather_object`` pickles through the host: ~1 ms at 8 ranks, a
    quarter of the sharded configs[3] sweep)."""
    import torch
    dist = _dist()
    per_rank = (n_rx + world - 1) // world
    width = 4 * n_freq + 3
    buf = np.zeros((per_rank, width))
    if part is not None:
        m = len(part["n_valid"])
        buf[:m, 0:n_freq] = part["fwhm_arcmin"]
        buf[:m, n_freq:2 * n_freq] = part["peak"]
        buf[:m, 2 * n_freq:4 * n_freq] = part["peak_index"].reshape(m, 2 * n_freq)  # < 2^53: exact
        buf[:m, 4 * n_freq] = part["n_valid"]
        buf[:m, 4 * n_freq + 1] = part["n_binned"]
        buf[:m, 4 * n_freq + 2] = part["mean_opl"]
    mine = torch.from_numpy(buf).to("cuda")
    out = torch.empty((world, per_rank, width), dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(out, mine, group=group)
    h = out.cpu().numpy()
    merged = {"fwhm_arcmin": np.zeros((n_rx, n_freq)), "peak": np.zeros((n_rx, n_freq)),
              "peak_index": np.zeros((n_rx, n_freq, 2), dtype=np.int64), "n_valid": np.zeros(n_rx),
              "n_binned": np.zeros(n_rx), "mean_opl": np.zeros(n_rx)}
    for r in range(world):
        idx = partition_items(n_rx, world, r)
        rows = h[r, :len(idx)]
        merged["fwhm_arcmin"][idx] = rows[:, 0:n_freq]
        merged["peak"][idx] = rows[:, n_freq:2 * n_freq]
        merged["peak_index"][idx] = np.rint(rows[:, 2 * n_freq:4 * n_freq]).astype(np.int64).reshape(-1, n_freq, 2)
        merged["n_valid"][idx] = rows[:, 4 * n_freq]
        merged["n_binned"][idx] = rows[:, 4 * n_freq + 1]
        merged["mean_opl"][idx] = rows[:, 4 * n_freq + 2]
    return merged


class SharedHostArray:
    """A host array in POSIX shared memory, mapped and page-locked (``cudaHostRegister``) by every
    rank of the node, so that each rank can DMA its part of a result straight into ONE array that
    the user's process sees whole.  Collective constructor / ``close``; one node only."""

    def __init__(self, shape, dtype=np.complex64, group=None):
        import torch
        from multiprocessing import shared_memory
        self.group = group
        self.rank, self.world = world_info(group)
        self.shape, self.dtype = tuple(int(v) for v in shape), np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        names = [None]
        if self.rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
            names = [self.shm.name]
        if self.world > 1:
            _dist().broadcast_object_list(names, src=0, group=group)
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=names[0])
            try:  # only the creating rank owns the segment (Python < 3.13 tracks attachments too)
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:  # noqa: BLE001
                pass
        self.array = np.ndarray((int(np.prod(self.shape)),), dtype=self.dtype, buffer=self.shm.buf)
        self.tensor = torch.from_numpy(self.array)
        self._registered = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), nbytes, 0)
            self._registered = int(rc) == 0
        if self.world > 1:
            _dist().barrier(group=group)

    def view(self):
        """The whole array in its logical shape (valid after ``PeerPlanes.gather_to_host``)."""
        return self.array.reshape(self.shape)

    def close(self):
        import torch
        if self.world > 1:
            _dist().barrier(group=self.group)
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self._registered = False
        self.tensor = None
        self.array = None
        try:
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except (BufferError, FileNotFoundError):
            pass
