"""The exchange step of a multi-GPU render behind the C ABI (include/lumen_b200.h lmb_comm_*, csrc/comm.cu): one NCCL all-reduce of
the LMB_FILM_SUM films with the "/ valid-sample count" epilogue queued behind it. On a one-GPU box the communicator has one rank
(the whole code path runs: NCCL kernel, resolve, snapshot + overlapped copy); with two GPUs the sharded render is held to the
single-GPU film of the same frames."""
import ctypes as C

import numpy as np
import pytest

from conftest import scene_path
from helpers import bits_equal
from lumen_b200 import host, integrator
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


def each(devs, fn):
    """one host thread per rank, as PathB200Multi does: a collective is entered by all ranks at once"""
    import threading
    errs = []

    def run(r, d):
        try:
            fn(r, d)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=run, args=(r, d)) for r, d in enumerate(devs)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        raise errs[0]


def test_single_rank_allreduce_resolves_in_place_and_overlapped():
    sc = host.Scene(scene_path("cornell"), 96, 64)
    pc, ubo = sc.make_pc(6, True), sc.make_ubo()
    dev = integrator.Device(0)
    try:
        dev.upload_scene(sc.desc)
        dev.build_accel()
        dev.init(96, 64, 4)
        assert dev.comm_info()[1] == 0
        with pytest.raises(RuntimeError, match="lmb_comm_init first"):
            dev.film_allreduce()
        dev.comm_init(integrator.comm_unique_id(), 0, 1)
        rank, size, version = dev.comm_info()
        assert (rank, size) == (0, 1) and version >= 21800
        dev.render(pc, ubo, 0, 8, 1, integrator.FILM_SUM)
        summed = dev.download()
        assert (summed[..., 3] == 8).all()
        # overlapped form: snapshot reduced + resolved on the comm stream into `out`, film cleared for the next batch
        out = np.full((64, 96, 4), -1.0, dtype=np.float32)
        dev.film_allreduce(out.ctypes.data, clear_film=True)
        dev.render(pc, ubo, 8, 8, 1, integrator.FILM_SUM)  # the next batch renders while the reduce runs
        dev.sync()
        want = summed.copy()
        want[..., :3] /= want[..., 3:4]
        want[..., 3] = 1.0
        assert bits_equal(out, want).all()
        second = dev.download()
        assert (second[..., 3] == 8).all() and not bits_equal(second, summed).all()  # the film was cleared and holds frames 8..15 only
        # in-place form: the film itself becomes the resolved image
        dev.film_allreduce()
        dev.sync()
        got = dev.download()
        want2 = second.copy()
        want2[..., :3] /= want2[..., 3:4]
        want2[..., 3] = 1.0
        assert bits_equal(got, want2).all()
        with pytest.raises(RuntimeError, match="clear_film needs out_rgba"):
            dev.film_allreduce(None, clear_film=True)
        # one receiver among one rank: lmb_film_reduce(root 0) is the same operation
        dev.render(pc, ubo, 16, 4, 1, integrator.FILM_SUM)  # adds to the resolved film: alpha 1 + 4
        a = dev.download()
        dev.film_reduce(0)
        dev.sync()
        b = dev.download()
        assert np.array_equal(b[..., :3], a[..., :3] / a[..., 3:4]) and (b[..., 3] == 1).all()
        with pytest.raises(RuntimeError, match="clear_film needs out_rgba on the root"):
            dev.film_reduce(0, None, clear_film=True)
        with pytest.raises(RuntimeError, match="root must be a rank"):
            dev.film_reduce(-1)
        # the mean of frames 0..7 is what the reference's running mean gives, to rounding
        cpu, _ = po.OracleScene(sc).render(pc, ubo, 0, 8)
        assert np.allclose(out[..., :3], cpu[..., :3], rtol=3e-6, atol=1e-7)
        dev.comm_destroy()
        assert dev.comm_info()[1] == 0
    finally:
        dev.close()


@pytest.mark.skipif(n_gpus() < 2, reason="needs two GPUs")
def test_two_gpus_allreduce_equals_the_single_gpu_film():
    """N-GPU image == 1-GPU image. Sample-index shards (rank r renders frames r, r + 2, ...) and pixel
    shards (rank r renders rows r, r + 2, ...), each reduced by ONE ncclAllReduce; every rank ends with the same bytes."""
    sc = host.Scene(scene_path("cornell"), 128, 96)
    pc, ubo = sc.make_pc(6, True), sc.make_ubo()
    single = integrator.Device(0)
    devs = [integrator.Device(0), integrator.Device(1)]
    try:
        single.upload_scene(sc.desc)
        single.build_accel()
        single.init(128, 96, 4)
        single.render(pc, ubo, 0, 8)
        want = single.download()
        for d in devs:
            d.upload_scene(sc.desc)
            d.build_accel()
        integrator.comm_init_all(devs)
        with pytest.raises(RuntimeError, match="two contexts on one device"):
            integrator.comm_init_all([single, integrator.Device(0)])
        # (a) sample-index shards
        for r, d in enumerate(devs):
            d.init(128, 96, 4)
            d.render(pc, ubo, r, 4, 2, integrator.FILM_SUM)
        each(devs, lambda r, d: (d.film_allreduce(), d.sync()))
        films = [d.download() for d in devs]
        assert films[0].tobytes() == films[1].tobytes()
        assert (films[0][..., 3] == 1).all()
        # against the same eight samples summed on ONE device the two-rank film differs only in the association of eight positive
        # fp32 terms (a few ulp); against the reference's running mean (mix() per frame, path.rgen:104-108) by the rounding of eight
        # lerps -- measured 4e-6 relative on a handful of dark pixels, so that bound is the looser one
        single.init(128, 96, 4)
        single.render(pc, ubo, 0, 8, 1, integrator.FILM_SUM)
        single.resolve()
        assert np.allclose(films[0][..., :3], single.download()[..., :3], rtol=1e-6, atol=1e-9)
        assert np.allclose(films[0][..., :3], want[..., :3], rtol=2e-5, atol=1e-7)
        # (b) pixel shards x all frames: every pixel is owned by one rank, so the reduced film is the single-GPU SUM film bit for bit
        single.init(128, 96, 4)
        single.render(pc, ubo, 0, 8, 1, integrator.FILM_SUM)
        single.resolve()
        want_sum = single.download()
        outs = [np.zeros((96, 128, 4), dtype=np.float32) for _ in devs]
        for r, d in enumerate(devs):
            d.set_pixel_shard(r, 2)
            d.init(128, 96, 4)
            d.render(pc, ubo, 0, 8, 1, integrator.FILM_SUM)
        each(devs, lambda r, d: (d.film_allreduce(outs[r].ctypes.data, clear_film=True), d.sync()))
        assert outs[0].tobytes() == outs[1].tobytes() == want_sum.tobytes()
        # (c) one receiver (lmb_film_reduce, root 1): the root's image is the all-reduce's, the other rank keeps its own film and buffer
        for r, d in enumerate(devs):
            d.render(pc, ubo, 0, 8, 1, integrator.FILM_SUM)
        before0 = devs[0].download()
        root_out, other_out = np.zeros((96, 128, 4), dtype=np.float32), np.full((96, 128, 4), -7.0, dtype=np.float32)
        each(devs, lambda r, d: (d.film_reduce(1, (root_out if r == 1 else other_out).ctypes.data, clear_film=(r == 1)), d.sync()))
        assert root_out.tobytes() == want_sum.tobytes() and (other_out == -7.0).all()
        assert devs[0].download().tobytes() == before0.tobytes() and (devs[1].download() == 0).all()
        with pytest.raises(RuntimeError, match="root is not a rank"):
            devs[0].film_reduce(2)
    finally:
        single.close()
        for d in devs:
            d.close()
