#!/usr/bin/env python
"""Karras tree vs clustered (PLOC) tree as the traversal tree, per scene: surface-area cost ratio (what LMB_TREE=auto decides on),
wide nodes / triangle tests per ray and k_trace time of a few frames. Run on the GPU box: python tools/tree_choice.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scenes"))
from lumen_b200 import host, integrator  # noqa: E402
import gen_classroom_standin as gen  # noqa: E402

SCENES = [("cornell", os.path.join(ROOT, "scenes/cornell_box/cornell_box_path.json"), 512, 512, 6, 16),
          ("caustics", os.path.join(ROOT, "scenes/caustics.json"), 1280, 720, 12, 8),
          ("materials", os.path.join(ROOT, "scenes/material_test/materials.json"), 512, 512, 10, 16),
          ("classroom-standin", gen.generate(os.path.join(ROOT, "scenes", "_generated", "classroom_standin"))[0], 1920, 1080, 8, 4)]
rows = []
for name, path, W, H, depth, frames in SCENES:
    sc = host.Scene(path, W, H)
    pc, ubo = sc.make_pc(depth, True), sc.make_ubo()
    for tree in ("lbvh", "ploc", "auto"):
        if tree == "auto":
            os.environ.pop("LMB_TREE", None)
        else:
            os.environ["LMB_TREE"] = tree
        dev = integrator.Device(0)
        dev.upload_scene(sc.desc)
        dev.build_accel()
        b = dev.stats()
        dev.init(W, H, 0)
        dev.render(pc, ubo, 0, frames)
        dev.set_profile_stages(True)
        dev.reset_stats()
        dev.render(pc, ubo, 0, frames)
        s = dev.stats()
        rows.append(dict(scene=name, tree=tree, walked="ploc" if b.ploc_iterations else "karras", cost_ratio=b.tree_cost_ratio, wide_nodes=b.wide_nodes,
                         nodes_per_ray=s.nodes_visited / s.rays, tris_per_ray=s.tris_tested / s.rays, trace_ms=s.ms_extend, build_ms=b.ms_build_accel))
        print(json.dumps(rows[-1]), flush=True)
        dev.close()
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "tree_choice.json"), "w"), indent=1)
