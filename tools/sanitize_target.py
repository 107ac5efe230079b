#!/usr/bin/env python
"""Small renders for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): cornell (area light: probe rays, all three
ray kinds), the classroom stand-in path (sun + sky: k_miss), a pixel shard, sum film + resolve, async download, array queries.
    compute-sanitizer --tool racecheck python tools/sanitize_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from lumen_b200 import host, integrator  # noqa: E402
from helpers import random_rays  # noqa: E402

dev = integrator.Device(0)
for path, W, H, depth in ((os.path.join(ROOT, "scenes/cornell_box/cornell_box_path.json"), 48, 40, 6),
                          (os.path.join(ROOT, "scenes/cornell_box/cornell_box_dir.json"), 40, 32, 5)):
    sc = host.Scene(path, W, H)
    dev.upload_scene(sc.desc)
    dev.build_accel()
    dev.set_pixel_shard(0, 1)
    dev.init(W, H, 2)
    pc, ubo = sc.make_pc(depth, True), sc.make_ubo()
    dev.render(pc, ubo, 0, 3)
    img = dev.download()
    dev.clear_film()
    dev.render(pc, ubo, 0, 2, 1, integrator.FILM_SUM)
    dev.resolve()
    pinned = np.empty((H, W, 4), np.float32)
    dev.download_async(pinned.ctypes.data)
    dev.sync()
    dev.set_pixel_shard(1, 3)
    dev.init(W, H, 2)
    dev.render(pc, ubo, 0, 2)
    dev.download()
    rays = random_rays(np.random.default_rng(5), [-3, -1, -3], [3, 5, 3], 4096)
    dev.trace_closest(rays)
    dev.trace_any(rays)
    print(os.path.basename(path), "ok", float(img[..., :3].mean()), dev.stats().rays)
dev.close()
