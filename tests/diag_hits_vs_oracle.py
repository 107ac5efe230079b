"""GPU diagnostic: closest / any-hit queries of the CUDA path vs the oracle, with details of every mismatch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import scene_path
from helpers import random_rays, bits_equal
from lumen_b200 import host, integrator
from oracle import pyoracle as po

dev = integrator.Device(0)
for name in sys.argv[1:] or ["cornell", "caustics", "materials"]:
    sc = host.Scene(scene_path(name), 64, 64)
    orc = po.OracleScene(sc)
    dev.upload_scene(sc.desc); dev.build_accel()
    st = dev.stats()
    print(name, "tris", sc.info.n_triangles, "wide nodes", st.wide_nodes, "levels", st.wide_levels, "build ms", st.ms_build_accel, "wide ms", st.ms_build_wide)
    rng = np.random.default_rng(21)
    box = orc.lbvh()["aabb"][:6]
    rays = random_rays(rng, box[:3], box[3:], 300000)
    gh, (ch, _) = dev.trace_closest(rays), orc.trace_closest(rays)
    bad = np.nonzero(~(bits_equal(gh["prim"], ch["prim"]) & bits_equal(gh["t"], ch["t"])))[0]
    print("  closest mismatches:", bad.size)
    for i in bad[:12]:
        print("   ray", i, rays[i], "gpu", gh["prim"][i], gh["t"][i], "cpu", ch["prim"][i], ch["t"][i])
    rays[:, 3] = 0.0
    rays[:, 7] = rng.uniform(0.05, 6.0, rays.shape[0])
    ga, ca = dev.trace_any(rays), orc.trace_any(rays)[0]
    print("  any mismatches:", int((ga != ca).sum()), "gpu occluded", int(ga.sum()), "cpu occluded", int(ca.sum()))
