#!/usr/bin/env python
"""ncu target for config 5 as bench.py runs it: 2^24 incoherent closest-hit rays through the 10 M-triangle torus grid, ordered inside
the launch (sort_rays). Launch order of k_trace_array: 2 probe launches of the build, then 2 x the measured launch.
    ncu --set full --clock-control none --import-source on -k regex:k_trace_array -s 3 -c 1 -f -o gpurun_out/prof_torus_sorted python tools/ncu_torus_sorted_target.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scenes"))
import torch  # noqa: E402
import gen_torus_grid as gen  # noqa: E402
from lumen_b200 import integrator  # noqa: E402

scene = gen.make_scene(10, 100, 50, 64, 64)
dev = integrator.Device(0)
dev.upload_scene(scene.desc)
dev.build_accel()
n = 1 << 24
d_rays = torch.from_numpy(gen.random_rays(n, 10)).cuda()
d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
for _ in range(2):
    ms = dev.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), 1, sort_rays=True)
    print("ms", ms, "Mrays/s", n / ms / 1e3)
