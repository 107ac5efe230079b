"""Small fixed workload for ncu captures: classroom stand-in 1920x1080, depth 8, `frames` frames per batch, `reps` batches.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lumen_b200 import integrator  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
scene = bench.load_scene()
dev = integrator.Device(0)
dev.upload_scene(scene.desc)
dev.build_accel()
dev.init(bench.WIDTH, bench.HEIGHT, frames)
pc, ubo = scene.make_pc(bench.MAX_DEPTH, True), scene.make_ubo()
for r in range(reps):
    dev.clear_film()
    dev.render(pc, ubo, r * frames, frames, 1, integrator.FILM_SUM)
st = dev.stats()
print("rays", st.rays, "ms_render", st.ms_render, "Mrays/s", st.rays / st.ms_render / 1e3, "launches", st.kernel_launches)
