"""N > 1 host logic on CPU: two gloo ranks shard the frame indices, accumulate sum films (samples come from the oracle,
standing in for the GPU renderer), all-reduce and resolve; the result must equal the single-process render."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import free_port, scene_path
from lumen_b200 import sharding


def _worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lumen_b200 import host
    from oracle import pyoracle as po
    sc = host.Scene(scene_path("cornell"), 40, 40)
    orc = po.OracleScene(sc)
    pc, ubo = sc.make_pc(6, True), sc.make_ubo()
    film = np.zeros((40, 40, 4), dtype=np.float32)
    frames = sharding.shard_frame_list(0, 3, rank, world)
    for f in frames:
        rgb, _ = orc.render_frame_raw(pc, ubo, f, threads=1)
        sharding.accumulate_sum(film, rgb)
    t = torch.from_numpy(film)
    sharding.all_reduce_film(t, dist)
    np.save(os.path.join(tmpdir, f"rank{rank}.npy"), sharding.resolve(t.numpy()))
    np.save(os.path.join(tmpdir, f"frames{rank}.npy"), np.array(frames))
    dist.destroy_process_group()


def test_two_rank_sharded_render_equals_single(tmp_path):
    world, port = 2, free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert r0.tobytes() == r1.tobytes()  # all-reduce leaves the same film on every rank
    f0, f1 = np.load(tmp_path / "frames0.npy").tolist(), np.load(tmp_path / "frames1.npy").tolist()
    assert sorted(f0 + f1) == list(range(6)) and not set(f0) & set(f1)
    from lumen_b200 import host
    from oracle import pyoracle as po
    sc = host.Scene(scene_path("cornell"), 40, 40)
    single, _ = po.OracleScene(sc).render(sc.make_pc(6, True), sc.make_ubo(), 0, 6)
    # running mean in frame order vs sum-then-divide: O(1e-7) relative (SURVEY.md 8e parity note)
    assert np.allclose(r0[..., :3], single[..., :3], rtol=3e-6, atol=1e-7)


def test_frame_partition_properties():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            seen += sharding.shard_frame_list(100, 5, r, world)
        assert sorted(seen) == list(range(100, 100 + 5 * world))


def test_resolve_handles_nan_and_empty_pixels():
    film = np.zeros((2, 2, 4), dtype=np.float32)
    s = np.ones((2, 2, 3), dtype=np.float32)
    s[0, 0] = np.nan  # NaN sample is skipped, count stays 0 -> pixel resolves to 0 (reference keeps the stale value)
    sharding.accumulate_sum(film, s)
    sharding.accumulate_sum(film, 3 * np.ones((2, 2, 3), dtype=np.float32))
    out = sharding.resolve(film)
    assert out[0, 0, :3].tolist() == [3, 3, 3] and out[1, 1, :3].tolist() == [2, 2, 2] and (out[..., 3] == 1).all()


def _tile_worker(rank, world, port, tmpdir):
    """Config-4 layout: 2 pixel shards (interleaved rows) x 1 sample shard; every rank writes only its rows of a full film."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lumen_b200 import host
    from oracle import pyoracle as po
    sc = host.Scene(scene_path("cornell"), 32, 30)
    orc = po.OracleScene(sc)
    pc, ubo = sc.make_pc(6, True), sc.make_ubo()
    p, s, P, S = sharding.grid_of_rank(rank, world, pixel_shards=2)
    rows = sharding.rows_of_shard(30, p, P)
    film = np.zeros((30, 32, 4), dtype=np.float32)
    for f in sharding.shard_frame_list(0, 2, s, S):
        rgb, _ = orc.render_frame_raw(pc, ubo, f, threads=1)  # the oracle renders whole frames; the shard keeps its rows
        part = np.zeros_like(film)
        sharding.accumulate_sum(part, rgb)
        film[rows] += part[rows]
    t = torch.from_numpy(film)
    sharding.all_reduce_film(t, dist)
    np.save(os.path.join(tmpdir, f"tile{rank}.npy"), sharding.resolve(t.numpy()))
    dist.destroy_process_group()


def test_two_rank_pixel_sharded_render_equals_single(tmp_path):
    world, port = 2, free_port()
    mp.spawn(_tile_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "tile0.npy"), np.load(tmp_path / "tile1.npy")
    assert r0.tobytes() == r1.tobytes()
    from lumen_b200 import host
    from oracle import pyoracle as po
    sc = host.Scene(scene_path("cornell"), 32, 30)
    single, _ = po.OracleScene(sc).render(sc.make_pc(6, True), sc.make_ubo(), 0, 2)
    assert np.allclose(r0[..., :3], single[..., :3], rtol=3e-6, atol=1e-7)


def test_pixel_x_sample_grid_partition():
    """Every (row, frame) pair of the job is rendered by exactly one rank for every P x S factorisation of 8 ranks."""
    height, frames_per_rank, world = 37, 3, 8
    for P in (1, 2, 4, 8):
        seen = {}
        for rank in range(world):
            p, s, P_, S = sharding.grid_of_rank(rank, world, P)
            assert P_ * S == world
            for y in sharding.rows_of_shard(height, p, P_):
                for f in sharding.shard_frame_list(0, frames_per_rank, s, S):
                    assert (int(y), f) not in seen
                    seen[(int(y), f)] = rank
        assert len(seen) == height * frames_per_rank * (world // P)
    try:
        sharding.grid_of_rank(0, 8, 3)
        assert False
    except ValueError:
        pass


def _bdpt_worker(rank, world, port, tmpdir):
    """The BDPT row shards the same way: a frame's sample is col + splat of that frame (quirk B2 makes the
    light-tracer image a function of the frame alone), so sample-index shards add up exactly like the Path integrator's."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lumen_b200 import host
    from lumen_b200._ctypes_types import PCBdpt
    from oracle import pyoracle as po
    sc = host.Scene(scene_path("cornell"), 32, 32)
    orc = po.OracleScene(sc)
    pc, ubo = PCBdpt.from_path_pc(sc.make_pc(5, True)), sc.make_ubo()
    film = np.zeros((32, 32, 4), dtype=np.float32)
    for f in sharding.shard_frame_list(0, 3, rank, world):
        col, splat, _ = orc.render_bdpt_frame_raw(pc, ubo, f, threads=1)
        sharding.accumulate_sum(film, col + splat)
    t = torch.from_numpy(film)
    sharding.all_reduce_film(t, dist)
    np.save(os.path.join(tmpdir, f"bdpt{rank}.npy"), sharding.resolve(t.numpy()))
    dist.destroy_process_group()


def test_two_rank_sharded_bdpt_equals_single(tmp_path):
    world, port = 2, free_port()
    mp.spawn(_bdpt_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "bdpt0.npy"), np.load(tmp_path / "bdpt1.npy")
    assert r0.tobytes() == r1.tobytes()
    from lumen_b200 import host
    from lumen_b200._ctypes_types import PCBdpt
    from oracle import pyoracle as po
    sc = host.Scene(scene_path("cornell"), 32, 32)
    single, _ = po.OracleScene(sc).render_bdpt(PCBdpt.from_path_pc(sc.make_pc(5, True)), sc.make_ubo(), 0, 6)
    assert np.allclose(r0[..., :3], single[..., :3], rtol=3e-6, atol=1e-7)
