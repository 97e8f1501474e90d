"""N>1 host logic on CPU: world_size-2 gloo processes shard a frame by row tiles, each fills
its tile (the CPU oracle stands in for the GPU kernel here — test-side use of the checker), one
gather assembles the frame, and the result equals the single-process render bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib

    import oracle_lib as O

    sh = importlib.import_module("raytracing-in-one-weekend_b200.sharding")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        W, H, spp = 48, 27, 4
        scene = O.rtb.host.make_scene("three_spheres")
        if mode == "equal":
            H = 28
            tiles = sh.row_tiles(H, world)
        else:
            cost = np.linspace(1.0, 4.0, H)
            tiles = sh.balanced_row_tiles(cost, world)
            assert tiles[0][1] - tiles[0][0] != tiles[1][1] - tiles[1][0]
        b, e = tiles[rank]
        p = O.rtb.host.make_params(scene, W, H, spp, 8, row_begin=b, row_end=e)
        buf = O.Buffers(W, H)
        O.sample_batch(scene, p, buf, threads=2)
        frame = torch.from_numpy(buf.out_color).reshape(H, W, 4)
        inside = frame[b:e].clone()
        assert float(frame[:b].abs().sum()) == 0.0 and float(frame[e:].abs().sum()) == 0.0   # only own rows written
        # the batched gather to rank 0 of several buffers at once (what bench.py times) ...
        extra = torch.from_numpy(buf.out_normal).reshape(H, W, 3).clone()
        to_root = [frame.clone(), extra]
        sh.gather_frames_to_root(to_root, tiles, root=0)
        if rank != 0:
            assert float(to_root[0][:b].abs().sum()) == 0.0 and float(to_root[0][e:].abs().sum()) == 0.0   # untouched off-root
        np.save(os.path.join(out_dir, f"root_{mode}_{rank}.npy"), to_root[0].numpy())
        np.save(os.path.join(out_dir, f"rootn_{mode}_{rank}.npy"), to_root[1].numpy())
        # ... and the all-gather form
        sh.gather_frame(frame, tiles)
        assert torch.equal(frame[b:e], inside)
        t = sh.max_over_ranks(float(rank + 1), "cpu")
        assert t == float(world)
        np.save(os.path.join(out_dir, f"frame_{mode}_{rank}.npy"), frame.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["equal", "balanced"])
def test_row_tile_shard_and_gather_world2(tmp_path, oracle, rtb, mode):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), mode, str(tmp_path)), nprocs=world, join=True)
    W, H, spp = 48, (28 if mode == "equal" else 27), 4
    scene = rtb.host.make_scene("three_spheres")
    p = rtb.host.make_params(scene, W, H, spp, 8)
    full = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, full)
    want = full.out_color.reshape(H, W, 4)
    for r in range(world):
        got = np.load(tmp_path / f"frame_{mode}_{r}.npy")
        assert np.array_equal(got, want)
    assert np.array_equal(np.load(tmp_path / f"root_{mode}_0.npy"), want)
    assert np.array_equal(np.load(tmp_path / f"rootn_{mode}_0.npy"), full.out_normal.reshape(H, W, 3))


def _shared_frame(W, H, rank, abi, rtb, stem):
    """HostBuffers over files in /dev/shm mapped by every rank: rank 0 creates them, the others map the same pages — the CPU
    stand-in for the frame that lives in rank 0's HBM and is mapped by the other ranks (renderer.FrameRenderer.enable_peer_frame)."""
    n = W * H
    spec = [("in_color", (n, 4), np.float32), ("in_weight", (n,), np.float32), ("in_normal", (n, 3), np.float32),
            ("in_albedo", (n, 3), np.float32), ("out_color", (n, 4), np.float32), ("out_weight", (n,), np.float32),
            ("out_normal", (n, 3), np.float32), ("out_albedo", (n, 3), np.float32), ("diagnostics", (n,), abi.DIAGNOSTICS_DTYPE)]
    if rank == 0:
        for name, shape, dt in spec:
            m = np.memmap(stem + name, dtype=dt, mode="w+", shape=shape)
            m[...] = 0
            m.flush()
    dist.barrier()
    hb = rtb.plugin.HostBuffers(1, 1)
    hb.width, hb.height = W, H
    for name, shape, dt in spec:
        setattr(hb, name, np.memmap(stem + name, dtype=dt, mode="r+", shape=shape))
    dist.barrier()
    if rank == 0:       # every rank holds its mapping now: the names can go
        for name, _, _ in spec:
            os.unlink(stem + name)
    return hb


def _shared_frame_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib

    import oracle_lib as O

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        W, H, spp = 40, 22, 4
        hb = _shared_frame(W, H, rank, O.abi, O.rtb, f"/dev/shm/rtb_test_{port}_")     # ONE frame mapped by both rank processes
        assert hb.out_color.shape == (W * H, 4) and hb.diagnostics.shape == (W * H,)
        scene = O.rtb.host.make_scene("three_spheres")
        # rank 0's cost model -> every rank (bench.py broadcasts the probe's row costs the same way), tiles from the
        # plugin's own balancer (rtb_balance_rows)
        cost = torch.from_numpy(np.linspace(1.0, 5.0, H)) if rank == 0 else torch.zeros(H, dtype=torch.float64)
        dist.broadcast(cost, src=0)
        bounds = O.rtb.plugin.balance_rows(cost.numpy(), 0, H, world)
        b, e = bounds[rank], bounds[rank + 1]
        p = O.rtb.host.make_params(scene, W, H, spp, 8, row_begin=b, row_end=e)
        O.sample_batch(scene, p, hb, threads=2)                        # each rank writes its own rows, in place
        done = torch.zeros(1)
        dist.all_reduce(done)                                          # the frame-complete signal (4 bytes on the GPU path)
        if rank == 0:                                                  # ... and rank 0 sees the whole frame without a gather
            np.save(os.path.join(out_dir, "shared_color.npy"), np.array(hb.out_color))
            np.save(os.path.join(out_dir, "shared_rays.npy"), np.array(hb.diagnostics["ray_count"]))
            np.save(os.path.join(out_dir, "bounds.npy"), np.array(bounds))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_one_frame_written_in_place_by_the_ranks_world2(tmp_path, oracle, rtb):
    """bench.py's N > 1 protocol on CPU: ONE frame every rank maps (rank 0's HBM over CUDA IPC on the GPU box, /dev/shm here),
    row tiles from the plugin's balancer on a cost model rank 0 broadcasts, each rank renders its tile into the frame in place,
    one tiny all-reduce marks the frame complete — no gather — and rank 0 holds the same bits as a single-process render."""
    world = 2
    mp.spawn(_shared_frame_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    W, H, spp = 40, 22, 4
    scene = rtb.host.make_scene("three_spheres")
    full = oracle.Buffers(W, H)
    oracle.sample_batch(scene, rtb.host.make_params(scene, W, H, spp, 8), full)
    assert np.array_equal(np.load(tmp_path / "shared_color.npy"), full.out_color)
    assert np.array_equal(np.load(tmp_path / "shared_rays.npy"), full.diagnostics["ray_count"])
    bounds = np.load(tmp_path / "bounds.npy")
    assert bounds[0] == 0 and bounds[-1] == H and bounds[1] > H // 2          # the cheap rows make the first tile the longer one
    assert not [f for f in os.listdir("/dev/shm") if f.startswith("rtb_test_")]      # the names are unlinked once mapped
