"""Host-side multi-GPU logic (vulkan_compute_ray_tracing_b200/sharding.py) on CPU: world_size-2/3 gloo process groups.

Each rank renders its share with the CPU oracle (the checker standing in for the CUDA library, which cannot run here)
through the product's shard_params / reduce / gather code; rank 0 compares the combined frame with the oracle's
unsharded render.  Tile sharding must be bit-exact, sample slicing equal within fp32 reassociation (rtol 1e-6).
"""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from vulkan_compute_ray_tracing_b200 import sharding  # noqa: E402

CAM = (1.8, 8.6, 1.1)
W, H = 200, 150     # ragged: 7 x 5 tiles, right/bottom tiles partly outside
SPP = 5             # not divisible by 2 or 3


def test_sample_slices_cover_exactly():
    for total in (1, 5, 64, 1024):
        for world in (1, 2, 3, 8):
            sl = sharding.sample_slices(total, world, first_sample=7)
            assert sl[0][0] == 7 and sum(n for _, n in sl) == total
            for (b0, n0), (b1, _) in zip(sl, sl[1:]):
                assert b1 == b0 + n0
            assert max(n for _, n in sl) - min(n for _, n in sl) <= 1


def test_tile_partition_and_packed_layout():
    for (w, h) in ((W, H), (1920, 1080), (33, 1), (32, 32)):
        for world in (1, 2, 3, 8):
            seen = np.zeros(w * h, np.int32)
            n = sharding.max_owned_tiles(w, h, world)
            for r in range(world):
                idx = sharding.tile_pixel_index(w, h, r, world, pad_tiles=n)
                assert idx.shape == (n * 1024,)
                seen[idx[idx >= 0]] += 1
                tiles = sharding.owned_tiles(w, h, r, world)
                assert np.all(tiles % world == r)
            assert seen.min() == 1 and seen.max() == 1


def test_shard_params_fields():
    from vulkan_compute_ray_tracing_b200._native import RenderParams
    p = RenderParams()
    p.sample_begin, p.sample_count = 10, 5
    q, active = sharding.shard_params(p, "tiles", 1, 4)
    assert (q.tile_rank, q.tile_count, q.sample_begin, q.sample_count, active) == (1, 4, 10, 5, True)
    got = [sharding.shard_params(p, "samples", r, 3) for r in range(3)]
    assert [(g.sample_begin, g.sample_count, a) for g, a in got] == [(10, 2, True), (12, 2, True), (14, 1, True)]
    # more ranks than samples: the surplus ranks render nothing
    assert [sharding.shard_params(p, "samples", r, 8)[1] for r in range(8)] == [True] * 5 + [False] * 3
    assert (p.tile_count, p.sample_count) == (0, 5)     # the caller's params are untouched
    with pytest.raises(ValueError):
        sharding.shard_params(p, "rows", 0, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, mode, result):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import GOLDEN
        from oracleharness import Oracle, make_params
        from refharness import load_scene
        scene = load_scene(os.path.join(GOLDEN, "doge_scene.vcrt"))
        o = Oracle()
        kw = dict(shader="full", max_bounces=4, accum="f32", rng="philox", trig="portable", philox_seed=3)
        base = make_params(sample_begin=2, sample_count=SPP, **kw)

        def render(p):
            acc = np.zeros((H, W, 4), np.float32)
            if p is not None:
                o.render(scene, CAM, W, H, p, accumf=acc)
            return torch.from_numpy(acc)

        if mode in ("tiles", "samples"):
            acc = sharding.render_sharded(render, base, mode, rank, world)
            got = acc.numpy()
        else:   # "gather": packed tiles -> all_gather -> unpack
            p, _ = sharding.shard_params(base, "tiles", rank, world)
            mine = render(p).reshape(-1, 4)
            n = sharding.max_owned_tiles(W, H, world)
            idx = torch.from_numpy(sharding.tile_pixel_index(W, H, rank, world, pad_tiles=n))
            packed = torch.zeros((n * 1024, 4), dtype=torch.float32)
            packed[idx >= 0] = mine[idx[idx >= 0]]
            out = torch.zeros((W * H, 4), dtype=torch.float32)
            got = sharding.gather_tiles(packed, W, H, world, out).reshape(H, W, 4).numpy()
        if rank == 0 or mode == "gather":
            want = o.render(scene, CAM, W, H, make_params(sample_begin=2, sample_count=SPP, **kw))["accumf"]
            if mode == "samples":
                ok = bool(np.allclose(got, want, rtol=1e-6, atol=1e-6)) and bool(np.array_equal(got[..., 3], want[..., 3]))
            else:
                ok = bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))
            result[rank] = ok
        else:
            result[rank] = True
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,world", [("tiles", 2), ("samples", 2), ("samples", 3), ("gather", 2), ("gather", 3)])
def test_sharded_render_matches_unsharded(oracle, mode, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        result = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, mode, result)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=180)
        for p in procs:
            if p.is_alive():
                p.terminate()
            assert p.exitcode == 0
        assert dict(result) == {r: True for r in range(world)}
