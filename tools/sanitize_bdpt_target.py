#!/usr/bin/env python
"""Small BDPT renders for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): all three pipelines (default pair-parallel,
LMB_BDPT=pixel, LMB_BDPT=mega) on cornell (area light), caustics (spot light, glass) and cornell_dir (directional light), running
mean and sum film.
    compute-sanitizer --tool memcheck python tools/sanitize_bdpt_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lumen_b200 import host, integrator  # noqa: E402
from lumen_b200._ctypes_types import PCBdpt  # noqa: E402

quick = "--quick" in sys.argv  # default pipeline, one scene (initcheck / racecheck are slow)
dev = integrator.Device(0)
for mode in (("pairs",) if quick else ("pairs", "pixel", "mega")):
    os.environ["LMB_BDPT"] = mode
    for rel, W, H, depth in (("scenes/cornell_box/cornell_box_path.json", 40, 32, 5), ("scenes/caustics.json", 32, 24, 7),
                             ("scenes/cornell_box/cornell_box_dir.json", 32, 32, 4))[:1 if quick else 3]:
        sc = host.Scene(os.path.join(ROOT, rel), W, H)
        dev.upload_scene(sc.desc)
        dev.build_accel()
        dev.init(W, H, 1)
        pc, ubo = PCBdpt.from_path_pc(sc.make_pc(depth, True), 3), sc.make_ubo()
        dev.render_bdpt(pc, ubo, 0, 2)
        img = dev.download()
        dev.clear_film()
        dev.render_bdpt(pc, ubo, 0, 2, 2, integrator.FILM_SUM)
        dev.resolve()
        print(mode, os.path.basename(rel), "ok", float(img[..., :3].mean()), dev.stats().rays)
dev.close()
