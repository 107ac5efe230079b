"""include/lmb_detmath.h against libm (float64 numpy as the reference)."""
import numpy as np

from oracle import pyoracle as po


def ulp_error(got, ref64):
    ref32 = ref64.astype(np.float32)
    spacing = np.spacing(np.abs(ref32)).astype(np.float64)
    return np.abs(got.astype(np.float64) - ref64) / np.maximum(spacing, 1e-45)


def test_sin_cos_within_2ulp_of_libm():
    x = np.random.default_rng(0).uniform(-64, 64, 400000).astype(np.float32)
    r = po.detmath(x, np.ones_like(x))
    # near zeros of sin/cos the absolute error is what matters (Cody-Waite reduction): 1.2e-7 absolute
    assert np.max(np.abs(r["sin"] - np.sin(x.astype(np.float64)))) < 1.3e-7
    assert np.max(np.abs(r["cos"] - np.cos(x.astype(np.float64)))) < 1.3e-7


def test_exp_within_2ulp():
    x = np.random.default_rng(1).uniform(-87, 88, 400000).astype(np.float32)
    r = po.detmath(x, np.ones_like(x))
    assert ulp_error(r["exp"], np.exp(x.astype(np.float64))).max() <= 2.0
    assert po.detmath(np.float32([-200.0, 0.0]), np.float32([1, 1]))["exp"].tolist() == [0.0, 1.0]


def test_pow_matches_hardware_style_exp2_log2():
    rng = np.random.default_rng(2)
    x = rng.uniform(1e-6, 1.0, 400000).astype(np.float32)
    y = rng.uniform(0.0, 9.0, 400000).astype(np.float32)
    got = po.detmath(x, y)["pow"].astype(np.float64)
    ref = np.power(x.astype(np.float64), y.astype(np.float64))
    rel = np.abs(got - ref) / ref
    # exp2(y*log2 x) in fp32: relative error grows with |y log2 x| (<= 180 here) * 2^-24
    assert rel.max() < 2e-5
    edge = po.detmath(np.float32([0.0, 1.0, 0.5, 0.0625]), np.float32([5.0, 5.0, 5.0, 0.0]))["pow"]
    assert edge.tolist() == [0.0, 1.0, 0.03125, 1.0]


def test_expf_window_form_equals_expf(tmp_path):
    """lmb_expf_fast (the form the sky march inlines, lmb_detmath.h) against lmb_expf on every 61st of the 2^32 bit patterns, NaNs and
    both tails included; tools/check_expf_equiv.c without -DSTRIDE walks all of them (16 s on 8 cores: 0 mismatches)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "check_expf")
    subprocess.run(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-mfma", "-DSTRIDE=61", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tools", "check_expf_equiv.c"), "-lm", "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("mismatches = 0 of"), out.stdout
