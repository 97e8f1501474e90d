"""FrameRenderer — the stand-in host's per-GPU driver around the plugin: device-resident
accumulation buffers and row-tile sharding over the ranks of one box.  One process per GPU
(torch.distributed / NCCL); with world size 1 it is just the device-buffer path of the C ABI.

Two ways to assemble the frame of N ranks:
  * peer frame (`enable_peer_frame`, the default of bench.py): the accumulation buffers live in rank 0's
    HBM (rtb_device_alloc), every other rank maps them (rtb_ipc_export / rtb_ipc_open) and its kernel reads
    and writes its row tile there over NVLink — the tiles land where the consumer reads them, nothing is
    gathered; one 4-byte all-reduce per frame tells rank 0's stream that every tile has landed;
  * gather (fallback, and what round 1 shipped): each rank renders into its own full-frame buffers and one
    batched NCCL send/recv group moves the tiles to rank 0.
(A host that drives all GPUs from ONE process needs neither: rtb_multi_sample_batch[_device].)

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
        self.peer = None              # peer-frame state: {"ptrs": {...}, "owned": bool}
        self._flag = None
        self._struct()

    def _struct(self):
        if self.peer is not None:
            p = self.peer["ptrs"]
            i, o = ("in_", "out_") if not self.peer.get("swapped") else ("out_", "in_")
            self.buffers = _plugin.device_buffers_struct(
                p[i + "color"], p[i + "weight"], p[i + "normal"], p[i + "albedo"],
                p[o + "color"], p[o + "weight"], p[o + "normal"], p[o + "albedo"], p.get("diag"))
            return
        self.buffers = _plugin.device_buffers_struct(
            self.inp["color"], self.inp["weight"], self.inp["normal"], self.inp["albedo"],
            self.out["color"], self.out["weight"], self.out["normal"], self.out["albedo"], self.diag)

    # ---- peer frame: rank 0 owns the buffers, the others write into them over NVLink -----------------
    class _Raw:                      # a raw device pointer as a __cuda_array_interface__ object
        def __init__(self, ptr, shape):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (int(ptr), False), "version": 2}

    def enable_peer_frame(self):
        """Moves the accumulation buffers into ONE allocation set on rank 0 that every rank maps.  Returns True when
        every rank succeeded (otherwise nothing changes and the gather path stays in use)."""
        if self.world == 1:
            return False
        n = self.width * self.height
        names = [("in_" + k, _ELEMS[k]) for k in ("color", "weight", "normal", "albedo")] + \
                [("out_" + k, _ELEMS[k]) for k in ("color", "weight", "normal", "albedo")]
        if self.diag is not None:
            names.append(("diag", 4))
        ok, ptrs, handles = True, {}, [None]
        try:
            if self.rank == 0:
                ptrs = {name: self.ctx.device_alloc(n * c * 4) for name, c in names}
                handles = [{name: self.ctx.ipc_export(p) for name, p in ptrs.items()}]
        except _plugin.RtbError:
            ok = False
        dist.broadcast_object_list(handles, src=0, group=self.group)
        if self.rank != 0:
            try:
                if handles[0] is None:
                    raise _plugin.RtbError(-1, "rank 0 could not export the frame")
                ptrs = {name: self.ctx.ipc_open(h) for name, h in handles[0].items()}
            except _plugin.RtbError:
                ok = False
        flag = torch.tensor([1.0 if ok else 0.0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if flag.item() < 1.0:
            for name, p in ptrs.items():
                try:
                    (self.ctx.device_free if self.rank == 0 else self.ctx.ipc_close)(p)
                except _plugin.RtbError:
                    pass
            return False
        self.peer = {"ptrs": ptrs, "elems": dict(names)}
        if self.rank == 0:           # rank 0 reads the frame through ordinary tensors over the same memory
            for k in ("color", "weight", "normal", "albedo"):
                self.inp[k] = torch.as_tensor(self._Raw(ptrs["in_" + k], (n, _ELEMS[k])), device=self.device)
                self.out[k] = torch.as_tensor(self._Raw(ptrs["out_" + k], (n, _ELEMS[k])), device=self.device)
            if self.diag is not None:
                self.diag = torch.as_tensor(self._Raw(ptrs["diag"], (n, 4)), device=self.device)
        self._flag = torch.zeros(1, device=self.device)
        self._struct()
        return True

    def frame_complete(self):
        """Peer frame: one 4-byte all-reduce, enqueued after this rank's kernel — when a rank's stream has passed it, every
        rank's tile has landed in rank 0's buffers.  Gather path: the gather itself."""
        if self.world == 1:
            return
        if self.peer is not None:
            dist.all_reduce(self._flag, group=self.group)
        else:
            self.gather()

    def set_tiles(self, tiles):
        self.tiles = tiles

    @property
    def my_rows(self):
        return self.tiles[self.rank]

    def swap(self):
        """accumulation := output (Raytracer.cs:798-802)."""
        self.inp, self.out = self.out, self.inp
        if self.peer is not None:
            self.peer["swapped"] = not self.peer.get("swapped", False)
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
        if self.peer is not None:
            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier(group=self.group)       # nobody unmaps or frees while a peer may still write
            for p in self.peer["ptrs"].values():
                try:
                    (self.ctx.device_free if self.rank == 0 else self.ctx.ipc_close)(p)
                except _plugin.RtbError:
                    pass
            self.peer = None
        self.ctx.close()
