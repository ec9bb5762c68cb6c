"""Sharding of the fit over the GPUs of one box; one process per GPU, torch.distributed for the plumbing.

Two axes, both used by `bench.py` and `FitSession`:

* frames  — every frame owns its activations w_f and pose (t_f, q_f); D, topology, UVs and cameras are replicated
            constants.  Frame ranges are independent units: NO data-path collective (BASELINE config 4).
* cameras — each rank renders a contiguous subset of the views of the SAME frames (`FitConfig.cam_slice`, optionally cut at
            bin-row granularity with `FitConfig.cam_band` so that 9 views balance over 2/4/8 ranks), the packed
            gradient vector [d_w | d_t | d_q] ((B+7) floats per frame) is all-reduced (sum) once per iteration and the
            Adam step is replicated (BASELINE config 5).  The loss of a view is scaled by 1 / C_total on every rank, so
            the partial gradients simply add up.

The reference itself is single-process (fit.py has no distributed code); this module is the multi-GPU part of the
north-star, not a restatement of reference code.
"""
import torch
import torch.distributed as dist


def split_range(n_items, rank, world):
    """Contiguous balanced split of range(n_items) over `world` ranks: the first n_items % world ranks get one extra
    item.  Returns (start, stop); empty when there are more ranks than items."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError('rank %r out of range for world size %r' % (rank, world))
    if n_items < 0:
        raise ValueError('n_items must be >= 0')
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def frame_shard(n_frames, rank=None, world=None):
    """Frames [start, stop) fitted by this rank (frame-sharded mode: no exchange between ranks)."""
    rank, world = _rank_world(rank, world)
    return split_range(n_frames, rank, world)


def camera_shard(n_cams, rank=None, world=None):
    """Views [start, stop) rendered by this rank in the camera-split mode (9 views over 2/4/8 ranks: 5+4, 3+2+2+2,
    2+1x7).  Pass the result as FitConfig.cam_slice."""
    rank, world = _rank_world(rank, world)
    s = split_range(n_cams, rank, world)
    if s[0] == s[1]:
        raise ValueError('camera split needs at least one view per rank (%d views, %d ranks)' % (n_cams, world))
    return s


def view_band_shard(n_cams, height, rank=None, world=None, bin_px=32):
    """Camera split at the granularity of bin rows (9 views do not divide evenly over 2/4/8 GPUs, a rank that
    renders 2 views while the others render 1 sets the pace).  The n_cams * R rows of 32-px bins (R = ceil(height / 32)) of
    all views, in (view, row) order, are split evenly: returns ((c0, c1), (row_lo, row_hi)) — this rank renders views
    [c0, c1), of view c0 only the bin rows >= row_lo and of view c1 - 1 only the bin rows < row_hi.  Pass them as
    FitConfig.cam_slice and FitConfig.cam_band.  Every pixel (and every antialias pixel pair, owned by its lower / left
    pixel) belongs to exactly one bin, so the partial losses and gradients of the ranks add up to the full ones."""
    rank, world = _rank_world(rank, world)
    R = -(-int(height) // bin_px)
    a, b = split_range(n_cams * R, rank, world)
    if a == b:
        raise ValueError('band split needs at least one bin row per rank (%d rows, %d ranks)' % (n_cams * R, world))
    return (a // R, (b - 1) // R + 1), (a % R, (b - 1) % R + 1)


def _rank_world(rank, world):
    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return 0, 1
    return rank, world


def allreduce_gradients(grads):
    """Sum the packed gradient vector over ranks in place (the only exchange of the camera-split mode).  NCCL over
    NVLink on GPU tensors, gloo on CPU tensors (tests); a no-op in a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(grads, op=dist.ReduceOp.SUM)
    return grads


def gather_frames(local, n_frames, dst=0):
    """Collect per-frame results [F_local, ...] of the frame-sharded fit on rank `dst` in frame order -> [n_frames, ...]
    on dst, None elsewhere.  (Role of fit.py:642's `result[frame] = ...` across ranks.)"""
    rank, world = _rank_world(None, None)
    if world == 1:
        return local
    sizes = [split_range(n_frames, r, world) for r in range(world)]
    parts = [torch.empty((b - a,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) for a, b in sizes] if rank == dst else None
    # point-to-point: the shards may have unequal sizes
    if rank == dst:
        for r in range(world):
            if r == dst:
                parts[r].copy_(local)
            elif parts[r].numel():
                dist.recv(parts[r], src=r)
    elif local.numel():
        dist.send(local.contiguous(), dst=dst)
    return torch.cat(parts) if rank == dst else None
