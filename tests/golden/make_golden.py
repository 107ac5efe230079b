"""Mints tests/golden/oracle_golden.npz from the oracle (run from the repo root: python tests/golden/make_golden.py).

The reference has no tests, golden images or known-answer vectors for the Path integrator and cannot be run
here (GLSL + Vulkan RT pipeline), so these vectors pin the ORACLE, not the reference: they detect accidental changes of the
restatement. Inputs are seeded; outputs are exact fp32 bit patterns."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lumen_b200 import host  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

rng = np.random.default_rng(20261017)
out = {}
out["pcg_in"] = rng.integers(0, 2**32, size=(64, 4), dtype=np.uint32)
out["pcg_in"][:4] = [[0, 0, 0, 0], [1, 2, 3, 4], [511, 511, 15, 1], [2**32 - 1] * 4]
out["pcg_out"] = po.pcg4d(out["pcg_in"])
out["rand_seed"] = np.array([[0, 0, 0, 0], [17, 4, 2, 0], [1919, 1079, 1023, 0]], dtype=np.uint32)
out["rand_out"] = po.rand(out["rand_seed"], 12)
sc = host.Scene(os.path.join(ROOT, "scenes/cornell_box/cornell_box_path.json"), 48, 48)
orc = po.OracleScene(sc)
img, st = orc.render(sc.make_pc(6, True), sc.make_ubo(), 0, 2)
out["cornell_48_d6_f2"] = img
out["cornell_48_d6_f2_rays"] = np.uint64(st.rays)
b = orc.lbvh()
out["cornell_left"], out["cornell_right"] = b["left"], b["right"]
sc2 = host.Scene(os.path.join(ROOT, "scenes/material_test/materials.json"), 40, 40)
img2, _ = po.OracleScene(sc2).render(sc2.make_pc(10, True), sc2.make_ubo(), 0, 2)
out["materials_40_d10_f2"] = img2
np.savez_compressed(os.path.join(ROOT, "tests/golden/oracle_golden.npz"), **out)
print("wrote", {k: getattr(v, "shape", None) for k, v in out.items()})
