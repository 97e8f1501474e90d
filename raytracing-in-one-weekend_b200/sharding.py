"""Row-tile sharding of one frame across the GPUs of a box.

Every pixel of the sample job is independent (SampleBatchJob.cs:72-76,159-163 touch only
element `index`) and the Philox stream is keyed by the GLOBAL pixel index, so any partition of
the rows renders the same image bit for bit.  Rank g renders rows [begin_g, end_g) straight
into its slice of a full-frame device buffer (rtb_batch_params.row_begin/row_end), and ONE
collective per frame — an in-place all-gather of the row tiles over NCCL/NVLink — assembles
the frame on every rank (`gather_frame`), or — what a host that reads the frame on one rank needs,
and what `bench.py` times — one batched gather of every buffer's tiles to rank 0
(`gather_frames_to_root`).  No other data-path communication exists.

The reference's own row mechanism (SliceOffset/SliceDivider, SampleBatchJob.cs:69) shards rows
in time; `interlaced_rows` exposes the same rule for callers that prefer row % N == g.
"""
import numpy as np
import torch
import torch.distributed as dist


def row_tiles(height, world):
    """Contiguous, near-equal row ranges: [(begin, end)] * world."""
    return [((g * height) // world, ((g + 1) * height) // world) for g in range(world)]


def balanced_row_tiles(row_cost, world):
    """Row ranges with near-equal total cost.  `row_cost`: per-row work estimate — the
    per-row sum of Diagnostics.RayCount of an earlier batch is the natural one (the reference
    keeps it for its MRays/s metric, Raytracer.cs:527-543)."""
    cost = np.asarray(row_cost, dtype=np.float64)
    height = len(cost)
    cum = np.concatenate([[0.0], np.cumsum(np.maximum(cost, 1e-12))])
    bounds = [0]
    for g in range(1, world):
        target = cum[-1] * g / world
        b = int(np.searchsorted(cum, target))
        b = min(max(b, bounds[-1] + 1), height - (world - g))
        bounds.append(b)
    bounds.append(height)
    return [(bounds[g], bounds[g + 1]) for g in range(world)]


def interlaced_rows(rank, world):
    """(SliceOffset, SliceDivider) giving rank `rank` the rows with row % world == rank."""
    return rank, world


def gather_frame(frame, tiles, group=None):
    """In-place gather of row tiles: on entry rank g holds valid rows tiles[g] of `frame`
    ([H, W, C] or [H*W, C] with H rows first); on return every rank holds the whole frame.

    Equal tiles -> one all_gather_into_tensor whose input is the rank's own slice of the output
    (no staging copy); unequal tiles -> one broadcast per tile (still one logical exchange)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return frame
    height = tiles[-1][1]
    rows = frame.reshape(height, -1)
    sizes = {e - b for b, e in tiles}
    if len(sizes) == 1 and tiles[0][0] == 0:
        b, e = tiles[rank]
        try:
            dist.all_gather_into_tensor(rows, rows[b:e], group=group)
            return frame
        except (RuntimeError, NotImplementedError):  # backend without the fused form
            pass
    works = []
    for g, (b, e) in enumerate(tiles):
        if e > b:
            works.append(dist.broadcast(rows[b:e], src=dist.get_global_rank(group, g) if group else g, group=group, async_op=True))
    for w in works:
        w.wait()
    return frame


def gather_frames_to_root(frames, tiles, root=0, group=None):
    """THE frame-end exchange: rank g sends its row tile of every buffer in `frames` (a list of
    [H*W, C] / [H, W, C] tensors) to `root`, as ONE batched group of point-to-point operations
    (a single ncclGroup: one launch whatever the number of buffers and ranks, equal tiles or not).
    On return `root` holds the whole frame in every buffer; other ranks are unchanged."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return frames
    height = tiles[-1][1]

    def glob(r):
        return dist.get_global_rank(group, r) if group is not None else r

    ops = []
    for f in frames:
        rows = f.reshape(height, -1)
        if rank == root:
            for g, (b, e) in enumerate(tiles):
                if g != root and e > b:
                    ops.append(dist.P2POp(dist.irecv, rows[b:e], glob(g), group))
        else:
            b, e = tiles[rank]
            if e > b:
                ops.append(dist.P2POp(dist.isend, rows[b:e], glob(root), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return frames


def max_over_ranks(value, device, group=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
