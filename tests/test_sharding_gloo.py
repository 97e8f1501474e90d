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


def _shared_frame_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib

    import oracle_lib as O

    bench = importlib.import_module("bench")
    sh = importlib.import_module("raytracing-in-one-weekend_b200.sharding")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        W, H, spp = 40, 22, 4
        hb = bench._shared_host_frame(W, H, rank, O.abi, O.rtb)       # ONE frame mapped by both rank processes
        assert hb is not None and hb.out_color.shape == (W * H, 4) and hb.diagnostics.shape == (W * H,)
        scene = O.rtb.host.make_scene("three_spheres")
        b, e = sh.row_tiles(H, world)[rank]
        p = O.rtb.host.make_params(scene, W, H, spp, 8, row_begin=b, row_end=e)
        O.sample_batch(scene, p, hb, threads=2)                        # each rank writes its own rows, in place
        dist.barrier()
        if rank == 0:                                                  # ... and rank 0 sees the whole frame without a gather
            np.save(os.path.join(out_dir, "shared_color.npy"), np.array(hb.out_color))
            np.save(os.path.join(out_dir, "shared_rays.npy"), np.array(hb.diagnostics["ray_count"]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_one_host_frame_shared_by_the_ranks_world2(tmp_path, oracle, rtb):
    """bench.py's N > 1 end-to-end path: the ranks map ONE set of host arrays (/dev/shm) and each renders its row tile into
    them in place; the frame is complete on the host without any exchange."""
    world = 2
    mp.spawn(_shared_frame_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    W, H, spp = 40, 22, 4
    scene = rtb.host.make_scene("three_spheres")
    full = oracle.Buffers(W, H)
    oracle.sample_batch(scene, rtb.host.make_params(scene, W, H, spp, 8), full)
    assert np.array_equal(np.load(tmp_path / "shared_color.npy"), full.out_color)
    assert np.array_equal(np.load(tmp_path / "shared_rays.npy"), full.diagnostics["ray_count"])
    assert not [f for f in os.listdir("/dev/shm") if f.startswith("rtb_bench_")]      # the names are unlinked once mapped
