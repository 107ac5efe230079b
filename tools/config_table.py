#!/usr/bin/env python
"""BASELINE.json configs 1-4 at their FULL sizes on one GPU (acceptance criterion 4): Mrays/s, spp/s, rays per
path, traversal work per ray and the roofline figure of k_trace.

    python tools/config_table.py [--configs 1,2,3,4] [--frames4 4096] [--out gpurun_out/config_table.json]

Product path only (the C-ABI library): parity of the same configs at full resolution against the oracle is
tests/test_gpu_full_configs.py, the CPU baseline is `bench.py --impl reference`. Config 5 (10 M-triangle torus grid) is
tools/torus_sweep.py.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scenes"))

import numpy as np  # noqa: E402

from lumen_b200 import host, integrator  # noqa: E402

NODE_BYTES, TRI_BYTES, RAY_HIT_BYTES = 80, 48, 48


def classroom_path():
    import gen_classroom_standin as gen
    path, _ = gen.generate(os.path.join(ROOT, "scenes", "_generated", "classroom_standin"))
    return path


CONFIGS = {
    # id: (label, scene path fn, W, H, frames, depth)
    1: ("cornell_box path, 512x512, 16 spp, depth 6", lambda: os.path.join(ROOT, "scenes/cornell_box/cornell_box_path.json"), 512, 512, 16, 6),
    2: ("caustics, 1280x720, 64 spp, depth 12", lambda: os.path.join(ROOT, "scenes/caustics.json"), 1280, 720, 64, 12),
    3: ("classroom stand-in, 1920x1080, 1024 spp, depth 8", classroom_path, 1920, 1080, 1024, 8),
    4: ("bedroom slot (classroom stand-in geometry, assets absent), 3840x2160, 4096 spp, depth 8", classroom_path, 3840, 2160, 4096, 8),
}


def run(cfg, dev, frames_override=None):
    label, path_fn, W, H, frames, depth = CONFIGS[cfg]
    if frames_override:
        frames = frames_override
    sc = host.Scene(path_fn(), W, H)
    dev.upload_scene(sc.desc)
    dev.build_accel()
    build = dev.stats()
    dev.init(W, H, 0)
    pc, ubo = sc.make_pc(depth, True), sc.make_ubo()
    dev.render(pc, ubo, 0, min(frames, 4))  # warm-up
    dev.reset_stats()
    dev.clear_film()
    t0 = time.perf_counter()
    dev.render(pc, ubo, 0, frames)  # synchronous on return
    wall = time.perf_counter() - t0
    st = dev.stats()
    film = dev.download()
    # stage split + traversal work of a profiled pass (CUDA events per stage)
    n_prof = min(frames, 8)
    dev.set_profile_stages(True)
    dev.reset_stats()
    dev.clear_film()
    dev.render(pc, ubo, 0, n_prof)
    ps = dev.stats()
    dev.set_profile_stages(False)
    trav_bytes = ps.nodes_visited * NODE_BYTES + ps.tris_tested * TRI_BYTES + ps.rays * RAY_HIT_BYTES
    row = {
        "config": cfg, "label": label, "triangles": int(sc.info.n_triangles), "width": W, "height": H, "frames": frames, "max_depth": depth,
        "rays": int(st.rays), "rays_closest": int(st.rays_closest), "rays_shadow": int(st.rays_shadow), "rays_probe": int(st.rays_probe),
        "ms_render_device": st.ms_render, "wall_s": wall, "mrays_per_s": st.rays / (st.ms_render * 1e-3) / 1e6, "spp_per_s": frames / (st.ms_render * 1e-3),
        "rays_per_path": st.rays / (frames * W * H), "kernel_launches": int(st.kernel_launches), "nan_samples": int(st.nan_samples),
        "nodes_per_ray": ps.nodes_visited / max(ps.rays, 1), "tris_per_ray": ps.tris_tested / max(ps.rays, 1),
        "stage_ms_per_frame": {k: v / n_prof for k, v in (("trace", ps.ms_extend), ("shade", ps.ms_shade), ("connect", ps.ms_connect), ("raygen_sky_film", ps.ms_film))},
        "k_trace_algorithmic_GBps": trav_bytes / (ps.ms_extend * 1e-3) / 1e9 if ps.ms_extend > 0 else 0.0,
        "accel_build_ms": build.ms_build_accel, "film_mean_rgb": [float(x) for x in film[..., :3].reshape(-1, 3).mean(0)],
    }
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4")
    ap.add_argument("--frames4", type=int, default=0, help="override config 4's frame count (default: the full 4096)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config_table.json"))
    a = ap.parse_args()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    dev = integrator.Device(0)
    rows = []
    for c in (int(x) for x in a.configs.split(",")):
        row = run(c, dev, a.frames4 if (c == 4 and a.frames4) else None)
        row["k_trace_frac_of_hbm_peak"] = row["k_trace_algorithmic_GBps"] / float(peaks.get("hbm_gbs", 6650.0))
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump({"hbm_peak_GBps": float(peaks.get("hbm_gbs", 6650.0)), "rows": rows}, open(a.out, "w"), indent=1)
    print("| config | Mrays/s | spp/s | rays/path | nodes/ray | tris/ray | k_trace GB/s (frac of HBM peak) | trace / shade / connect / sky+film ms per frame |")
    print("|---|---|---|---|---|---|---|---|")
    for r in rows:
        st = r["stage_ms_per_frame"]
        print(f"| {r['config']}: {r['label']} | {r['mrays_per_s']:.0f} | {r['spp_per_s']:.1f} | {r['rays_per_path']:.2f} | {r['nodes_per_ray']:.2f} | {r['tris_per_ray']:.2f} | "
              f"{r['k_trace_algorithmic_GBps']:.0f} ({r['k_trace_frac_of_hbm_peak']:.2f}) | {st['trace']:.2f} / {st['shade']:.2f} / {st['connect']:.2f} / {st['raygen_sky_film']:.2f} |")


if __name__ == "__main__":
    main()
