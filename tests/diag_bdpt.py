"""Diagnostic (not collected by pytest): GPU BDPT vs the oracle per scene -- fraction of bit-equal pixels, worst relative error,
ray counts, first mismatching pixel -- and device time per frame. Run on a GPU box: python tests/diag_bdpt.py [size]."""
import sys
import time

import numpy as np

from conftest import scene_path
from helpers import bits_equal, pixel_agreement
from lumen_b200 import host, integrator
from lumen_b200._ctypes_types import PCBdpt
from oracle import pyoracle as po

size = int(sys.argv[1]) if len(sys.argv) > 1 else 96
dev = integrator.Device(0)
for name, depth in [("cornell", 6), ("caustics", 8), ("materials", 7), ("cornell_dir", 5)]:
    sc = host.Scene(scene_path(name), size, size)
    dev.upload_scene(sc.desc)
    dev.build_accel()
    dev.init(size, size, 1)
    pc = PCBdpt.from_path_pc(sc.make_pc(depth, True))
    ubo = sc.make_ubo()
    osc = po.OracleScene(sc)
    dev.reset_stats()
    col, splat = dev.kat_bdpt_frame_raw(pc, ubo, 0)
    st = dev.stats()
    ocol, osplat, ost = osc.render_bdpt_frame_raw(pc, ubo, 0)
    same = bits_equal(col, ocol).all(axis=-1)
    rel = np.abs(col - ocol) / np.maximum(np.abs(ocol), 1e-6)
    print(f"{name}: col bit-equal {same.mean():.5f} within1e-4 {pixel_agreement(col, ocol):.5f} splat within1e-4 {pixel_agreement(splat, osplat):.5f} "
          f"rays gpu {st.rays_closest}/{st.rays_shadow} cpu {ost.rays_closest}/{ost.rays_shadow} ms {st.ms_render:.2f} "
          f"means {col.mean():.5f} {ocol.mean():.5f} splat {splat.mean():.5f} {osplat.mean():.5f}", flush=True)
    if not same.all():
        ys, xs = np.nonzero(~same)
        for y, x in list(zip(ys, xs))[:5]:
            print("   mismatch at", x, y, col[y, x], ocol[y, x], "rel", np.nanmax(rel[y, x]))
    osc.close()
# throughput at a larger size (device time, rays from the device counters)
import os
for mode in ("pairs", "pixel", "mega"):
  os.environ["LMB_BDPT"] = mode
  for name, depth, big in [("cornell", 6, 512)]:
      sc = host.Scene(scene_path(name), big, big)
      dev.upload_scene(sc.desc)
      dev.build_accel()
      dev.init(big, big, 1)
      pc = PCBdpt.from_path_pc(sc.make_pc(depth, True))
      ubo = sc.make_ubo()
      dev.render_bdpt(pc, ubo, 0, 2)
      dev.reset_stats()
      t0 = time.time()
      dev.render_bdpt(pc, ubo, 2, 8)
      st = dev.stats()
      rays = st.rays_closest + st.rays_shadow
      print(f"{mode} {name} {big}x{big} depth {depth}: {st.ms_render / 8:.2f} ms/frame, {rays / st.ms_render / 1e3:.1f} Mrays/s ({rays / 8 / big / big:.2f} rays/pixel), wall {time.time() - t0:.2f}s")
dev.close()
