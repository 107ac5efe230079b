"""lumen_headless: the C++ host shell (PathB200 : Integrator -> init / create_accel / render / update / destroy, scene ingest, EXR
output; lumen_b200/host/main.cpp, path_b200.h) driven end to end as a user would, against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, scene_path
from lumen_b200 import host

EXE = os.path.join(ROOT, "lumen_b200", "host", "lumen_headless")


def _run(args, **kw):
    return subprocess.run([EXE] + args, capture_output=True, text=True, timeout=600, **kw)


def test_cli_is_built_and_fails_loudly_without_a_device(tmp_path):
    """No CPU fallback: without CUDA the shell reports the C ABI's error and exits non-zero, writing nothing."""
    assert os.path.exists(EXE), "run __graft_entry__.build()"
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    out = tmp_path / "x.exr"
    r = _run([scene_path("cornell"), "--width", "16", "--height", "16", "--spp", "1", "--out", str(out)])
    assert r.returncode != 0 and "no CPU fallback" in r.stderr and not out.exists()
    r = _run([scene_path("cornell"), "--integrator", "bdpt", "--width", "16", "--height", "16", "--spp", "1", "--out", str(out)])  # BDPTB200, same rule
    assert r.returncode != 0 and "no CPU fallback" in r.stderr and not out.exists()


@pytest.mark.gpu
def test_cli_render_equals_oracle_at_half_precision(tmp_path):
    from oracle import pyoracle as po
    out = tmp_path / "cornell.exr"
    r = _run([scene_path("cornell"), "--width", "96", "--height", "80", "--spp", "6", "--depth", "6", "--batch", "4", "--out", str(out)])
    assert r.returncode == 0, r.stderr
    assert "wrote" in r.stdout and "Mrays/s" in r.stdout
    img = host.load_exr(str(out))
    sc = host.Scene(scene_path("cornell"), 96, 80)
    # --batch 4 with 6 spp renders frames 0..3 (whole batches only, like RayTracer's frame loop with a fixed batch)
    cpu, _ = po.OracleScene(sc).render(sc.make_pc(6, True), sc.make_ubo(), 0, 4)
    want = cpu[..., :3].astype(np.float16).astype(np.float32)  # save_exr stores HALF (ImageUtils.cpp:22-89)
    got = img[..., :3]
    assert got.shape == want.shape
    close = np.abs(got - want) <= 2.0 ** -10 * np.maximum(np.abs(want), 1e-4)
    assert close.mean() > 0.999


@pytest.mark.gpu
def test_cli_checkpoint_resume_and_rmse(tmp_path):
    """Progressive service: --ref / --target-rmse stop, --checkpoint / --resume continue bit-identically."""
    full, part, ck = tmp_path / "full.exr", tmp_path / "part.exr", tmp_path / "ck.bin"
    base = [scene_path("cornell"), "--width", "64", "--height", "64", "--depth", "5", "--batch", "2"]
    assert _run(base + ["--spp", "8", "--out", str(full)]).returncode == 0
    assert _run(base + ["--spp", "4", "--out", str(part), "--checkpoint", str(ck)]).returncode == 0
    r = _run(base + ["--spp", "8", "--out", str(part), "--checkpoint", str(ck), "--resume"])
    assert r.returncode == 0 and "resumed" in r.stdout
    assert host.load_exr(str(part)).tobytes() == host.load_exr(str(full)).tobytes()
    r = _run(base + ["--spp", "8", "--out", str(part), "--ref", str(full), "--target-rmse", "1e9"])
    assert r.returncode == 0 and "true RMSE" in r.stdout


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0,0", "0,1"])
def test_cli_multi_device_render_equals_single(tmp_path, devices):
    """--devices: PathB200Multi shards the sample index over the listed devices (one host thread and one context each), reduces
    the sum films device to device (lmb_film_add_from) and resolves. Same image as one device up to the order of fp32 adds."""
    if devices == "0,1" and _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    one, two = tmp_path / "one.exr", tmp_path / "two.exr"
    base = [scene_path("cornell"), "--width", "80", "--height", "64", "--depth", "6", "--spp", "8", "--batch", "2"]
    assert _run(base + ["--out", str(one)]).returncode == 0
    r = _run(base + ["--out", str(two), "--devices", devices])
    assert r.returncode == 0, r.stderr
    assert "8 frames on 2 devices" in r.stdout
    a, b = host.load_exr(str(one))[..., :3], host.load_exr(str(two))[..., :3]
    close = np.abs(a - b) <= 2.0 ** -9 * np.maximum(np.abs(a), 1e-4)  # one half ulp either way
    assert close.mean() > 0.999
    r = _run(base + ["--devices", devices, "--checkpoint", str(tmp_path / "c.bin")])
    assert r.returncode != 0 and "cannot be combined" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0,0", "0,1"])
def test_cli_multi_device_bdpt_equals_single(tmp_path, devices):
    """--integrator bdpt --devices: BDPTB200Multi (the multi-GPU host of the BDPT row, same sharding and reduce as PathB200Multi). With a
    fixed --time the sharded render is the single-device one up to the order of fp32 adds (light-tracer splats are float atomics)."""
    if devices == "0,1" and _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    one, two = tmp_path / "one.exr", tmp_path / "two.exr"
    base = [scene_path("cornell"), "--integrator", "bdpt", "--time", "7", "--width", "64", "--height", "48", "--depth", "5", "--spp", "8", "--batch", "2"]
    assert _run(base + ["--out", str(one)]).returncode == 0
    r = _run(base + ["--out", str(two), "--devices", devices])
    assert r.returncode == 0, r.stderr
    assert "8 frames on 2 devices, BDPT" in r.stdout and ("NCCL" in r.stdout) == (devices == "0,1")
    a, b = host.load_exr(str(one))[..., :3], host.load_exr(str(two))[..., :3]
    close = np.abs(a - b) <= 2.0 ** -8 * np.maximum(np.abs(a), 1e-3)
    assert close.mean() > 0.995
