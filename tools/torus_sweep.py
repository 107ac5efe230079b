#!/usr/bin/env python
"""BASELINE.json config 5: synthetic 10 M-triangle torus grid -- LBVH build time by stage and closest-hit Mrays/s for
(i) primary camera rays and (ii) incoherent rays (uniform origins / directions in the grid box), 2^20 .. 2^26 rays.

    python tools/torus_sweep.py [--K 10 --nu 100 --nv 50] [--max-log2 26] [--out gpurun_out/torus_sweep.json]

Size-independent checks at full size: the 8-wide tree passes the structural check (every triangle in exactly one leaf,
every child box contains its subtree) and the wide walk returns bit-identical hits to the binary LBVH walk on 2^20 rays.
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
import torch  # noqa: E402

import gen_torus_grid as gen  # noqa: E402
from lumen_b200 import integrator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=10)
    ap.add_argument("--nu", type=int, default=100)
    ap.add_argument("--nv", type=int, default=50)
    ap.add_argument("--max-log2", type=int, default=26)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "torus_sweep.json"))
    a = ap.parse_args()
    t0 = time.time()
    scene = gen.make_scene(a.K, a.nu, a.nv, 2048, 2048)
    t_gen = time.time() - t0
    dev = integrator.Device(0)
    dev.upload_scene(scene.desc)
    builds = []
    for _ in range(3):
        dev.build_accel()
        s = dev.stats()
        builds.append(dict(total=s.ms_build_accel, morton=s.ms_build_morton, sort=s.ms_build_sort, tree=s.ms_build_tree, refit_pack=s.ms_build_refit,
                           wide=s.ms_build_wide))
    st = dev.stats()
    chk = dev.wide_bvh_check()
    n_tris = int(scene.info.n_triangles)
    res = {"triangles": n_tris, "K": a.K, "nu": a.nu, "nv": a.nv, "host_generate_s": t_gen, "lbvh_build_ms": builds, "wide_nodes": int(st.wide_nodes),
           "wide_levels": int(st.wide_levels), "wide_check": chk, "sweeps": []}
    print(json.dumps({k: res[k] for k in ("triangles", "lbvh_build_ms", "wide_nodes", "wide_levels", "wide_check")}), flush=True)

    def run(kind, rays_np, sort_rays=False):
        n = rays_np.shape[0]
        d_rays = torch.from_numpy(rays_np).cuda()
        d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        dev.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), 1, sort_rays)  # warm-up
        dev.reset_stats()
        reps = 5 if n <= (1 << 24) else 3
        ms = dev.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), reps, sort_rays)
        s = dev.stats()
        hit_frac = float((d_hits[:, 3].view(torch.int32) != -1).float().mean().item())
        traced = s.rays_closest
        out = {"kind": kind, "rays": n, "ms_per_launch": ms / reps, "mrays_per_s": n * reps / ms / 1e3, "hit_fraction": hit_frac,
               "nodes_per_ray": s.nodes_visited / max(traced, 1), "tris_per_ray": s.tris_tested / max(traced, 1),
               "algorithmic_bytes_per_ray": (s.nodes_visited * 80 + s.tris_tested * 48) / max(traced, 1) + 48}
        out["achieved_GBps"] = out["algorithmic_bytes_per_ray"] * out["mrays_per_s"] / 1e3
        res["sweeps"].append(out)
        print(json.dumps(out), flush=True)
        return d_rays, d_hits

    # (i) primary camera rays of a 2048 x 2048 view (2^22 rays)
    ubo = scene.make_ubo()
    inv_view = np.array(ubo.inv_view, np.float32).reshape(4, 4).T  # column-major -> matrix
    inv_proj = np.array(ubo.inv_projection, np.float32).reshape(4, 4).T
    w = h = 2048
    px, py = np.meshgrid(np.arange(w, dtype=np.float32) + 0.5, np.arange(h, dtype=np.float32) + 0.5)
    ndc = np.stack([px / w * 2 - 1, py / h * 2 - 1, np.ones_like(px), np.ones_like(px)], -1).reshape(-1, 4)
    tgt = ndc @ inv_proj.T
    d = tgt[:, :3] / np.linalg.norm(tgt[:, :3], axis=1, keepdims=True)
    d = d @ inv_view[:3, :3].T
    org = np.broadcast_to(inv_view[:3, 3], d.shape)
    prim_rays = np.ascontiguousarray(np.concatenate([org, np.full((d.shape[0], 1), 1e-3), d, np.full((d.shape[0], 1), 1e4)], 1), dtype=np.float32)
    run("primary 2048x2048", prim_rays)
    # (ii) incoherent rays
    keep = None
    for lg in range(20, a.max_log2 + 1, 2):
        rays = gen.random_rays(1 << lg, a.K)
        r = run("incoherent", rays)
        if lg == 20:
            keep = (rays, r[1].cpu().numpy().copy())
        rs = run("incoherent, rays sorted by (origin cell, direction bin) inside the timed launch", rays, sort_rays=True)
        if not torch.equal(r[1].view(torch.int32), rs[1].view(torch.int32)):
            raise SystemExit("torus_sweep: sorted launch returned different hits")
    # wide walk == binary walk at full size (2^20 incoherent rays)
    os.environ["LMB_TRAVERSAL"] = "bvh2"
    dev2 = integrator.Device(0)
    dev2.upload_scene(scene.desc)
    dev2.build_accel()
    hb = dev2.trace_closest(keep[0])
    hw = keep[1]
    same = bool((hb["prim"] == hw[:, 3].view(np.uint32)).all() and (hb["t"].view(np.uint32) == hw[:, 0].view(np.uint32)).all()
                and (hb["b1"].view(np.uint32) == hw[:, 1].view(np.uint32)).all())
    res["wide_equals_binary_on_2^20_rays"] = same
    d_rays = torch.from_numpy(keep[0]).cuda()
    d_hits = torch.empty((keep[0].shape[0], 4), dtype=torch.float32, device="cuda")
    dev2.trace_closest_device(d_rays.data_ptr(), keep[0].shape[0], d_hits.data_ptr(), 1)
    ms = dev2.trace_closest_device(d_rays.data_ptr(), keep[0].shape[0], d_hits.data_ptr(), 3)
    res["binary_walk_mrays_per_s_2^20"] = keep[0].shape[0] * 3 / ms / 1e3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    res["hbm_peak_GBps"] = peaks.get("hbm_gbs", 6650.0)
    print(json.dumps({k: res[k] for k in ("wide_equals_binary_on_2^20_rays", "binary_walk_mrays_per_s_2^20", "hbm_peak_GBps")}), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    if not same or chk["errors"] or chk["dup_or_missing"]:
        raise SystemExit("torus_sweep: parity property violated")


if __name__ == "__main__":
    main()
