#!/usr/bin/env python
"""ncu target for the HBM-bound case of k_trace: 2^22 incoherent closest-hit rays through the 10 M-triangle torus grid
(BASELINE.json config 5). Launch order: build (incl. 2 probe launches of k_trace_array), then 3 x the measured launch.
    ncu --set full --clock-control none --import-source on -k regex:k_trace_array -s 3 -c 1 -f -o gpurun_out/prof_torus python tools/ncu_torus_target.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scenes"))
import torch  # noqa: E402
import gen_torus_grid as gen  # noqa: E402
from lumen_b200 import integrator  # noqa: E402

scene = gen.make_scene(10, 100, 50, 2048, 2048)
dev = integrator.Device(0)
dev.upload_scene(scene.desc)
dev.build_accel()
rays = gen.random_rays(1 << 22, 10)
d_rays = torch.from_numpy(rays).cuda()
d_hits = torch.empty((rays.shape[0], 4), dtype=torch.float32, device="cuda")
for _ in range(3):
    ms = dev.trace_closest_device(d_rays.data_ptr(), rays.shape[0], d_hits.data_ptr(), 1)
    print("ms", ms, "Mrays/s", rays.shape[0] / ms / 1e3)
