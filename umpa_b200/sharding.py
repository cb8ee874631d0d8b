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
        H = sam[0].shape[0]
        self.n_rows = H - 2 * self.padding
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
        if self.model is None:
            return {}
        if "abc" in kw and kw["abc"] is not None:
            kw["abc"] = kw["abc"][self.band[0]:self.band[1]]
        return self.model.match_device(**kw)

    def gather(self, local, keys=("f", "T", "dx", "dy", "df", "err"), dst=0):
        """Concatenate the bands' maps on rank `dst` (returns None elsewhere)."""
        return gather_bands(local, self.bands, self.rank, keys=keys, dst=dst)


def gather_bands(local, bands, rank, keys=("f", "T", "dx", "dy", "df", "err"), dst=0):
    """The only inter-GPU traffic of a sharded match: rank r contributes maps of shape
    (bands[r][1]-bands[r][0], N1, ...); rank `dst` receives their row-wise concatenation.
    One torch.distributed.gather per map (NCCL for CUDA tensors, gloo for CPU tensors)."""
    import torch.distributed as dist
    world = len(bands)
    out = {}
    for k in keys:
        t = local.get(k)
        meta = [None] * world
        dist.all_gather_object(meta, None if t is None else (tuple(t.shape[1:]), str(t.dtype).split(".")[-1]))
        known = [m for m in meta if m is not None]
        if not known:
            continue
        tail, dtype = known[0][0], getattr(torch, known[0][1])
        if t is None:               # a rank with an empty band still takes part
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
            t = torch.empty((0,) + tuple(tail), dtype=dtype, device=dev)
        # gather wants equal shapes: bands differ by at most one row, pad to the tallest
        rows = max(b[1] - b[0] for b in bands)
        send = t.contiguous()
        if send.shape[0] < rows:
            send = torch.cat([send, send.new_zeros((rows - send.shape[0],) + tuple(tail))], dim=0)
        parts = None
        if rank == dst:
            parts = [torch.empty((rows,) + tuple(tail), dtype=dtype, device=t.device) for _ in bands]
        dist.gather(send, parts, dst=dst)
        if rank == dst:
            out[k] = torch.cat([p_[:b[1] - b[0]] for p_, b in zip(parts, bands)], dim=0)
    return out if rank == dst else None
