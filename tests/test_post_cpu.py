"""Host-side steps either side of the integrator, no GPU: the oracle's restatement of the
EXR half conversion pinned against the REAL tinyexr the reference vendors (compiled into liblumen_host.so from
/root/reference/libs/tinyexr.h), the EXR writer fed by pre-converted half planes, and the progressive-render checkpoint."""
import os

import numpy as np
import pytest

from lumen_b200 import host
from oracle import pyoracle as po


def _special_floats(rng, n):
    """Values that exercise every branch of float_to_half_full: normals, ties, half subnormals, underflow, overflow,
    fp32 subnormals, signed zeros, infinities, NaNs."""
    x = np.concatenate([
        rng.normal(size=n).astype(np.float32) * 10.0,
        np.exp(rng.uniform(-30, 15, n)).astype(np.float32) * rng.choice([-1, 1], n).astype(np.float32),
        (rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)).view(np.float32),  # raw bit patterns (incl. NaN, inf, denormals)
        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 65504.0, 65519.9, 65520.0, 1e30, -1e30, 6.1035156e-05, 6.0e-05, 5.9604645e-08,
                  2.9802322e-08, 2.98e-08, 1e-10, 1e-45, 1.0, 1.0 + 2.0**-11, 1.0 + 2.0**-11 + 2.0**-23, 1.0 + 3 * 2.0**-11, 0.1, 1 / 3], dtype=np.float32),
    ])
    return x.astype(np.float32)


def test_half_conversion_pinned_to_tinyexr(tmp_path):
    """save_exr (real tinyexr, float in -> HALF out) followed by load_exr (half -> float) must give exactly the values of
    the oracle's float_to_half restatement, for every branch of the conversion."""
    rng = np.random.default_rng(5)
    vals = _special_floats(rng, 4000)
    w = 97
    h = (vals.size + 3 * w - 1) // (3 * w)
    img = np.zeros((h, w, 4), np.float32)
    flat = np.zeros(h * w * 3, np.float32)
    flat[: vals.size] = vals
    img[..., :3] = flat.reshape(h, w, 3)
    path = str(tmp_path / "a.exr")
    host.save_exr(img, path)
    back = host.load_exr(path)[..., :3]
    want = po.float_to_half(img[..., :3]).view(np.float16).astype(np.float32)
    same = (back.view(np.uint32) == want.view(np.uint32)) | (np.isnan(back) & np.isnan(want))
    assert same.all(), f"{(~same).sum()} of {same.size} values differ from tinyexr"
    # round-half-up, not ties-to-even: this is where numpy's float16 cast differs
    tie = np.array([1.0 + 2.0**-11], np.float32)
    assert po.float_to_half(tie)[0] == 0x3C01 and tie.astype(np.float16).view(np.uint16)[0] == 0x3C00


def test_exr_from_half_planes_is_byte_identical(tmp_path):
    rng = np.random.default_rng(6)
    img = np.zeros((33, 47, 4), np.float32)
    img[..., :3] = np.exp(rng.uniform(-12, 8, (33, 47, 3))).astype(np.float32)
    img[..., 3] = 1.0
    a, b = str(tmp_path / "float.exr"), str(tmp_path / "half.exr")
    host.save_exr(img, a)
    planes = np.stack([po.float_to_half(img[..., c]) for c in (2, 1, 0)])  # B, G, R
    host.save_exr_half_bgr(planes, b)
    assert open(a, "rb").read() == open(b, "rb").read()


def test_checkpoint_round_trip_and_rejects_garbage(tmp_path):
    rng = np.random.default_rng(7)
    film = rng.normal(size=(19, 31, 4)).astype(np.float32)
    film[3, 4, 0] = np.nan
    p = str(tmp_path / "run.ckpt")
    host.save_checkpoint(p, film, frames=123, path_length=6)
    back, frames, depth = host.load_checkpoint(p)
    assert frames == 123 and depth == 6 and back.shape == film.shape
    assert back.tobytes() == film.tobytes()
    assert not os.path.exists(p + ".tmp")
    blob = open(p, "rb").read()
    open(p, "wb").write(blob[:-5])  # truncated
    with pytest.raises(RuntimeError):
        host.load_checkpoint(p)
    open(p, "wb").write(b"\0" * 64)  # wrong magic
    with pytest.raises(RuntimeError):
        host.load_checkpoint(p)
    with pytest.raises(RuntimeError):
        host.load_checkpoint(str(tmp_path / "missing.ckpt"))
