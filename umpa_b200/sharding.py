"""Row-band sharding of one match() over the GPUs of a box.

Every output pixel depends only on input rows within +-padding of it, so the output rows are
split into contiguous bands, one per rank; rank r builds an ordinary model on the input rows
[r0, r1 + 2*padding) of every frame (band + halo) and runs the single-GPU kernels.  There is no
collective on the data path; the only traffic is the optional gather of the result maps."""
import numpy as np
import torch


def row_bands(n_rows, world):
    """Split n_rows output rows into `world` contiguous bands [(r0, r1), ...] (sizes differ by <= 1)."""
    base, extra = divmod(int(n_rows), int(world))
    bands, r = [], 0
    for k in range(world):
        n = base + (1 if k < extra else 0)
        bands.append((r, r + n))
        r += n
    return bands


def band_input_rows(band, padding):
    """Input rows of every frame that output rows [r0, r1) need."""
    r0, r1 = band
    return r0, r1 + 2 * padding


class ShardedMatcher:
    """One rank's part of a row-sharded match.

    cls: UMPAModelNoDF / UMPAModelDF / UMPAModelDFKernel; frames: full stacks (numpy or torch)
    that this rank can slice.  ``match_device`` returns the local band's maps; ``gather`` collects
    them on rank 0 with torch.distributed (NCCL on GPUs, gloo on CPU tensors)."""

    def __init__(self, cls, sam, ref, rank, world, window_size=2, max_shift=4, mask=None, **kw):
        self.rank, self.world = int(rank), int(world)
        self.padding = int(max_shift) + int(window_size) + cls.safe_crop
        if kw.get("pos_list") is not None or any(tuple(f.shape) != tuple(sam[0].shape) for f in sam):
            raise NotImplementedError("ShardedMatcher: equal frames at position 0 only (a ragged / stepped stack has no "
                                      "common row bands); match it on one GPU")
        self.masked = mask is not None
        H = sam[0].shape[0]
        self.n_rows = H - 2 * self.padding
        self.n_cols = sam[0].shape[1] - 2 * self.padding
        self.bands = row_bands(self.n_rows, self.world)
        self.band = self.bands[self.rank]
        lo, hi = band_input_rows(self.band, self.padding)
        take = lambda stack: None if stack is None else [f[lo:hi] for f in stack]
        contiguous = lambda fr: None if fr is None else [
            (f.contiguous() if isinstance(f, torch.Tensor) else np.ascontiguousarray(f)) for f in fr]
        self.model = None
        if self.band[1] > self.band[0]:
            self.model = cls(contiguous(take(sam)), contiguous(take(ref)), mask_list=contiguous(take(mask)),
                             window_size=window_size, max_shift=max_shift, **kw)

    def match_device(self, **kw):
        """The local band's maps (torch CUDA tensors).  `step` / `ROI` are not supported (the bands are cut in
        unstrided output rows); masked models take the coverage threshold of the WHOLE frame, as the unsharded
        match does (model.pyx:431): the bands' coverage maxima are all-reduced (MAX) first."""
        if kw.get("step") is not None or kw.get("ROI") is not None:
            raise NotImplementedError("ShardedMatcher.match_device: step / ROI are not supported")
        if self.masked:
            import torch.distributed as dist
            local = float(self.model.coverage().max()) if self.model is not None else 0.
            if self.world > 1 and dist.is_available() and dist.is_initialized():
                dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
                t = torch.tensor([local], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                local = float(t.item())
            kw["cover_max"] = local
        if self.model is None:
            return {}
        if "abc" in kw and kw["abc"] is not None:
            kw["abc"] = kw["abc"][self.band[0]:self.band[1]]
        return self.model.match_device(**kw)

    def gather(self, local, keys=("f", "T", "dx", "dy", "df", "err"), dst=0):
        """Concatenate the bands' maps on rank `dst` (returns None elsewhere)."""
        return gather_bands(local, self.bands, self.rank, keys=keys, dst=dst, n_cols=self.n_cols)


_INT_KEYS = ("err", "debug_Ncalls")


def gather_bands(local, bands, rank, keys=("f", "T", "dx", "dy", "df", "err"), dst=0, n_cols=None):
    """The only inter-GPU traffic of a sharded match (north_star: "a final gather of the output maps"): rank r
    contributes maps of shape (bands[r][1]-bands[r][0], N1); rank `dst` receives their row-wise concatenation.

    Point to point, one batch (torch.distributed.batch_isend_irecv -- one NCCL group over NVLink for CUDA tensors, gloo
    for CPU tensors), no Python objects exchanged: every rank sends each of its maps straight into the row slice it
    occupies in the full map on `dst`, nothing packed or copied twice.  With fewer than four ranks on NCCL a rank's
    float64 maps travel stacked as one message and its int32 maps as another (_gather_bands_packed): NCCL runs the
    messages to ONE peer one after the other, and seven 16 MB messages are slower than one of 100 MB (config 2 on two
    B200: 0.96 -> 0.59 ms on one box, 2.4 -> 1.05 ms on another), while with seven peers the per-map messages already
    run side by side (eight B200: 0.82 ms per map against 0.98 ms packed; config 4: 1.19 against 1.59 ms).  An
    all-gather of everything to everybody is not faster at either size (tools/diag_gather.py).  UMPA_GATHER=maps |
    packed | allgather forces a variant.  Every rank must pass the same `keys`; a key the model does not produce (df
    for NoDF) is skipped on every rank alike.  n_cols: map width
    (needed by a `dst` whose own band is empty; it then expects every key)."""
    import os
    import torch.distributed as dist
    n_rows = bands[-1][1]
    mine = bands[rank][1] - bands[rank][0]
    some = next((local[k] for k in keys if local.get(k) is not None), None)
    ops, out = [], {}
    mode = os.environ.get("UMPA_GATHER", "auto")
    if mode == "allgather":
        return _gather_bands_allgather(local, bands, rank, keys, dst, n_cols)
    if mode == "packed" or (mode == "auto" and len(bands) < 4 and dist.get_backend() == "nccl"):
        return _gather_bands_packed(local, bands, rank, keys, dst, n_cols)
    if rank == dst:
        if some is not None:
            n_cols, dev = int(some.shape[1]), some.device
        else:
            if n_cols is None:
                raise ValueError("gather_bands: a destination with an empty band needs n_cols")
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        for k in keys:
            t = local.get(k)
            if t is None and mine:
                continue
            full = torch.empty((n_rows, n_cols), dtype=torch.int32 if k in _INT_KEYS else torch.float64, device=dev)
            if mine:
                full[bands[rank][0]:bands[rank][1]] = t
            for r, (r0, r1) in enumerate(bands):
                if r != dst and r1 > r0:
                    ops.append(dist.P2POp(dist.irecv, full[r0:r1], r))
            out[k] = full
    elif mine:
        for k in keys:
            t = local.get(k)
            if t is not None:
                ops.append(dist.P2POp(dist.isend, t.contiguous(), dst))
    for req in (dist.batch_isend_irecv(ops) if ops else []):
        req.wait()
    return out if rank == dst else None


def _gather_bands_packed(local, bands, rank, keys, dst, n_cols):
    """gather_bands with TWO messages per peer (its float64 maps stacked, its int32 maps stacked) instead of one per
    map: with a single peer NCCL runs the seven 16 MB messages of config 2 one after the other (2.5 ms on two B200
    against 0.6 ms packed); with seven peers the per-map messages already run in parallel.  Every rank must hold the
    same set of keys (the maps of one model class)."""
    import torch.distributed as dist
    n_rows = bands[-1][1]
    mine = bands[rank][1] - bands[rank][0]
    k64 = [k for k in keys if k not in _INT_KEYS and (local.get(k) is not None or not mine)]
    k32 = [k for k in keys if k in _INT_KEYS and (local.get(k) is not None or not mine)]
    ops, out, staged = [], {}, []
    if rank == dst:
        some = next((local[k] for k in keys if local.get(k) is not None), None)
        if some is not None:
            n_cols, dev = int(some.shape[1]), some.device
        else:
            if n_cols is None:
                raise ValueError("gather_bands: a destination with an empty band needs n_cols")
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        for k in k64 + k32:
            out[k] = torch.empty((n_rows, n_cols), dtype=torch.int32 if k in _INT_KEYS else torch.float64, device=dev)
            if mine:
                out[k][bands[rank][0]:bands[rank][1]] = local[k]
        for r, (r0, r1) in enumerate(bands):
            if r == dst or r1 <= r0:
                continue
            for ks, dt in ((k64, torch.float64), (k32, torch.int32)):
                if ks:
                    buf = torch.empty((len(ks), r1 - r0, n_cols), dtype=dt, device=dev)
                    ops.append(dist.P2POp(dist.irecv, buf, r))
                    staged.append((ks, r0, r1, buf))
    elif mine:
        for ks in (k64, k32):
            if ks:
                ops.append(dist.P2POp(dist.isend, torch.stack([local[k] for k in ks]), dst))
    for req in (dist.batch_isend_irecv(ops) if ops else []):
        req.wait()
    for ks, r0, r1, buf in staged:
        for n, k in enumerate(ks):
            out[k][r0:r1] = buf[n]
    return out if rank == dst else None


def _gather_bands_allgather(local, bands, rank, keys, dst, n_cols):
    """gather_bands as ONE collective per dtype: every rank's maps, stacked and padded to the tallest band, go through
    all_gather_into_tensor (all NVLink channels, where point-to-point messages to one peer use a few); `dst` unpacks
    its copy, the other ranks drop theirs.  Needs every rank to hold the same keys and a non-empty band."""
    import torch.distributed as dist
    world, n_rows = len(bands), bands[-1][1]
    rows = max(r1 - r0 for r0, r1 in bands)
    some = next(local[k] for k in keys if local.get(k) is not None)
    n_cols, dev = int(some.shape[1]), some.device
    out = {}
    for ks, dt in (([k for k in keys if k not in _INT_KEYS and local.get(k) is not None], torch.float64),
                   ([k for k in keys if k in _INT_KEYS and local.get(k) is not None], torch.int32)):
        if not ks:
            continue
        mine = torch.empty((len(ks), rows, n_cols), dtype=dt, device=dev)
        for n, k in enumerate(ks):
            mine[n, :local[k].shape[0]] = local[k]
        allb = torch.empty((world * len(ks), rows, n_cols), dtype=dt, device=dev)     # (concatenated along dim 0)
        dist.all_gather_into_tensor(allb, mine)
        allb = allb.view(world, len(ks), rows, n_cols)
        if rank == dst:
            for n, k in enumerate(ks):
                full = torch.empty((n_rows, n_cols), dtype=dt, device=dev)
                for r, (r0, r1) in enumerate(bands):
                    full[r0:r1] = allb[r, n, :r1 - r0]
                out[k] = full
    return out if rank == dst else None
