"""Mints tests/golden/oracle_bdpt_golden.npz from the oracle's BDPT restatement (oracle/bdpt.h); run from the repo root:
python tests/golden/make_golden_bdpt.py. Like oracle_golden.npz these vectors pin the ORACLE (the reference ships none and its BDPT
is not even deterministic: oracle/bdpt.h B1-B2); they detect accidental changes of the restatement. Exact fp32 bit patterns."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lumen_b200 import host  # noqa: E402
from lumen_b200._ctypes_types import PCBdpt  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

out = {}
for name, path, size, depth in [("cornell", "scenes/cornell_box/cornell_box_path.json", 40, 5), ("materials", "scenes/material_test/materials.json", 32, 7),
                                ("caustics", "scenes/caustics.json", 32, 6), ("cornell_dir", "scenes/cornell_box/cornell_box_dir.json", 32, 4)]:
    sc = host.Scene(os.path.join(ROOT, path), size, size)
    orc = po.OracleScene(sc)
    pc = PCBdpt.from_path_pc(sc.make_pc(depth, True), 7)
    col, splat, st = orc.render_bdpt_frame_raw(pc, sc.make_ubo(), 2, threads=1)
    out[f"{name}_col"], out[f"{name}_splat"] = col, splat
    out[f"{name}_rays"] = np.array([st.rays_closest, st.rays_shadow], dtype=np.uint64)
    out[f"{name}_cfg"] = np.array([size, depth, 7, 2], dtype=np.int64)  # size, max_depth, time, frame
    orc.close()
np.savez_compressed(os.path.join(ROOT, "tests/golden/oracle_bdpt_golden.npz"), **out)
print("wrote", {k: getattr(v, "shape", None) for k, v in out.items()})
