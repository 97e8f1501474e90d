"""FrameRenderer — the stand-in host's per-GPU driver around the plugin: device-resident
accumulation buffers, row-tile sharding over the ranks of one box, and the one gather per
frame.  One process per GPU (torch.distributed / NCCL); with world size 1 it is just the
device-buffer path of the C ABI.

PyTorch is plumbing here (device memory, streams, NCCL); every pixel is produced by
rtb_sample_batch_device (librtb.so).
"""
import torch
import torch.distributed as dist

from . import _abi as abi
from . import plugin as _plugin
from . import sharding as _sharding

_ELEMS = {"color": 4, "weight": 1, "normal": 3, "albedo": 3, "diag": 4}


class FrameRenderer:
    def __init__(self, scene, width, height, device_index=0, group=None, tiles=None, diagnostics=True):
        self.width, self.height = width, height
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device("cuda", device_index)
        torch.cuda.set_device(self.device)
        self.ctx = _plugin.Context(device_index)
        self.ctx.upload(scene)
        self.tiles = tiles or _sharding.row_tiles(height, self.world)
        n = width * height

        def buf(c):
            return torch.zeros(n, c, device=self.device, dtype=torch.float32)

        # accumulation (in) and output sets, full frame on every rank; a rank only touches its rows
        self.inp = {k: buf(_ELEMS[k]) for k in ("color", "weight", "normal", "albedo")}
        self.out = {k: buf(_ELEMS[k]) for k in ("color", "weight", "normal", "albedo")}
        self.diag = buf(4) if diagnostics else None
        self._struct()

    def _struct(self):
        self.buffers = _plugin.device_buffers_struct(
            self.inp["color"], self.inp["weight"], self.inp["normal"], self.inp["albedo"],
            self.out["color"], self.out["weight"], self.out["normal"], self.out["albedo"], self.diag)

    def set_tiles(self, tiles):
        self.tiles = tiles

    @property
    def my_rows(self):
        return self.tiles[self.rank]

    def swap(self):
        """accumulation := output (Raytracer.cs:798-802)."""
        self.inp, self.out = self.out, self.inp
        self._struct()

    def clear(self):
        for t in self.inp.values():
            t.zero_()

    def _tile_params(self, params):
        p = abi.BatchParams.from_buffer_copy(params)
        p.row_begin, p.row_end = self.my_rows
        return p

    def render_device(self, params, gather=True, gather_aovs=True, all_ranks=False):
        """Enqueue one batch for this rank's row tile on torch's current stream, then the frame gather."""
        b, e = self.my_rows
        if e > b:
            self.ctx.sample_batch_device(self._tile_params(params), self.buffers, torch.cuda.current_stream().cuda_stream)
        if gather:
            self.gather(gather_aovs, all_ranks)

    def gather(self, aovs=True, all_ranks=False):
        """The one exchange per frame.  Default: every buffer's row tiles to rank 0 in ONE batched NCCL
        send/recv group (the host reads the frame on one rank).  all_ranks=True: in-place all-gather
        of each buffer instead, for consumers that need the frame everywhere."""
        if self.world == 1:
            return
        keys = ("color", "weight", "normal", "albedo") if aovs else ("color",)
        frames = [self.out[k] for k in keys]
        if self.diag is not None and aovs:
            frames.append(self.diag)
        if all_ranks:
            for f in frames:
                _sharding.gather_frame(f, self.tiles, self.group)
        else:
            _sharding.gather_frames_to_root(frames, self.tiles, 0, self.group)

    def render_host(self, params, host, fetch_all_ranks=False):
        """End-to-end batch with (pinned) HOST accumulators `host` (dict of torch CPU tensors with the
        keys of `inp`/`out`): H2D of this rank's rows, kernel, gather, D2H of the frame (rank 0, or every
        rank).  Returns (h2d_bytes, d2h_bytes) moved by this rank."""
        b, e = self.my_rows
        w = self.width
        h2d = d2h = 0
        for k in ("color", "weight", "normal", "albedo"):
            src = host["in_" + k].view(self.height * w, -1)[b * w:e * w]
            self.inp[k][b * w:e * w].copy_(src, non_blocking=True)
            h2d += src.numel() * 4
        self.render_device(params, all_ranks=fetch_all_ranks)
        if self.rank == 0 or fetch_all_ranks:
            for k in ("color", "weight", "normal", "albedo"):
                host["out_" + k].view(self.height * w, -1).copy_(self.out[k], non_blocking=True)
                d2h += self.out[k].numel() * 4
            if self.diag is not None and "diag" in host:
                host["diag"].view(self.height * w, -1).copy_(self.diag, non_blocking=True)
                d2h += self.diag.numel() * 4
        torch.cuda.current_stream().synchronize()
        return h2d, d2h

    def render_host_in_place(self, params, shared):
        """End-to-end batch on ONE set of pinned host arrays every rank maps (`shared`: plugin.HostBuffers over
        the same pages in all rank processes, registered with `ctx.register_host_buffers`): each rank's
        rtb_sample_batch reads and writes its own row tile of the host frame in place over its GPU's PCIe
        link, so the frame needs no gather.  Returns the (read, written) host bytes of this rank."""
        b, e = self.my_rows
        if e > b:
            self.ctx.sample_batch(self._tile_params(params), shared)
        px = (e - b) * self.width
        return px * 44, px * (48 + 16)

    def close(self):
        self.ctx.close()
