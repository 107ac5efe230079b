"""BDPT throughput on the product path (no oracle): device time per frame, Mrays/s and spp/s per scene. Run on a GPU box:
    python tools/bdpt_table.py [--json gpurun_out/bdpt_table.json] [--quick] [--scene NAME]
--quick renders one scene at 512 x 512 for ncu (tools: ncu -k regex:k_bdpt ... python tools/bdpt_table.py --quick)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lumen_b200 import host, integrator  # noqa: E402
from lumen_b200._ctypes_types import PCBdpt  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
quick = "--quick" in sys.argv
out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
cases = [("cornell", os.path.join(ROOT, "scenes/cornell_box/cornell_box_path.json"), 512, 512, 6, 16)]
if not quick:
    cases += [("caustics", os.path.join(ROOT, "scenes/caustics.json"), 1280, 720, 12, 8),
              ("materials", os.path.join(ROOT, "scenes/material_test/materials.json"), 1024, 1024, 10, 8),
              ("classroom_standin", None, bench.WIDTH, bench.HEIGHT, bench.MAX_DEPTH, 4)]
if "--scene" in sys.argv:  # one scene only (ncu launch lists)
    want = sys.argv[sys.argv.index("--scene") + 1]
    cases = [c for c in cases if c[0] == want]
dev = integrator.Device(0)
rows = []
for name, path, w, h, depth, frames in cases:
    sc = bench.load_scene() if path is None else host.Scene(path, w, h)
    dev.upload_scene(sc.desc)
    dev.build_accel()
    dev.init(w, h, 1)
    pc = PCBdpt.from_path_pc(sc.make_pc(depth, True))
    ubo = sc.make_ubo()
    dev.render_bdpt(pc, ubo, 0, 1 if quick else 2)  # warm-up (allocates the vertex buffers)
    dev.reset_stats()
    dev.render_bdpt(pc, ubo, 2, frames)
    st = dev.stats()
    rays = st.rays_closest + st.rays_shadow
    row = dict(scene=name, width=w, height=h, max_depth=depth, frames=frames, ms_per_frame=st.ms_render / frames, mrays_s=rays / st.ms_render / 1e3,
               spp_s=frames / (st.ms_render * 1e-3), rays_per_pixel=rays / frames / (w * h), nodes_per_ray=st.nodes_visited / max(rays, 1),
               tris_per_ray=st.tris_tested / max(rays, 1), nan_samples=st.nan_samples)
    rows.append(row)
    print(json.dumps(row), flush=True)
dev.close()
if out_json:
    with open(out_json, "w") as f:
        json.dump(rows, f, indent=1)
