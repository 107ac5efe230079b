#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 path integrator (BASELINE.json metric: Mrays/s and spp/s, classroom 1080p).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): the classroom configuration of BASELINE.json -- 1920x1080, max depth 8, sun + sky, Disney /
diffuse / conductor materials. The reference's classroom meshes are not in its tree (scenes/classroom/how-to-obtain.txt),
so the geometry is the labelled procedural stand-in of scenes/gen_classroom_standin.py, loaded through the same
Mitsuba-XML path. One STEP = one pass of the hot path over one batch: every rank renders `frames_per_step` frames
(samples per pixel) of the full image into its sum film; with N > 1 the films are summed on rank 0 by ONE NCCL reduce behind the C ABI
(lmb_film_reduce: snapshot -> ncclReduce -> "/ count" epilogue on the context's comm stream, overlapping the next step's
rendering) and every rank gets the resolved image. Ranks take disjoint frame indices (frame = first + rank + k*N): per-GPU
work is fixed -> weak scaling.

`value`  = traced rays of all ranks / time of K steps, scene + film resident in HBM, timed between barriers +
           torch.cuda.synchronize(), max over ranks.
`e2e`    = same metric through the public C-ABI calls with HOST buffers inside the timed region: per step the push
           constants + camera UBO are handed over from host memory (lmb_render copies them) and the resolved RGBA32F film lands in
           pinned host memory (lmb_download_async / lmb_film_reduce's out pointer on rank 0).
`roofline`        the dominant kernel, k_trace, on this workload: ISSUE bound (the BVH is L2-resident): warp instructions issued
                  (the kernel's own counters x the per-counter costs of profiles/ktrace_calibration.json) / time, against
                  SMs x 4 schedulers x the SM clock sampled in this run. The algorithmic-bytes figure is kept as a secondary field.
`roofline_config5` the same kernel on BASELINE config 5 (10 M-triangle torus grid, 2^24 incoherent rays: the HBM-sized case), measured in
                  this run (N = 1): algorithmic bytes / time against the measured HBM peak.
`config4`         BASELINE config 4's shape beside the headline: 3840x2160, 2 pixel shards x N/2 sample shards, ONE frame per rank and
                  step, reduced by the same call.
`film_check`      (N > 1) the N-rank resolved film against the same frames rendered by rank 0 alone.
`cpu_baseline`    the reference's own shader source compiled for the CPU (oracle/_ref/libglslref.so, kind "reference"; the oracle port
                  when that library is absent) on a bounded sample of the same workload at full resolution.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

WIDTH, HEIGHT, MAX_DEPTH = 1920, 1080, 8
WORKLOAD = "classroom-standin {w}x{h} depth 8 (procedural stand-in for scenes/classroom, whose meshes are not in the reference tree)"
NODE_BYTES, TRI_BYTES, RAY_BYTES, HIT_BYTES = 80, 48, 32, 16  # 8-wide compressed node, 3 x float4 triangle, ray in, hit out
SLOT_BYTES = 330  # wavefront state per path slot


def load_scene(width=None, height=None):
    sys.path.insert(0, os.path.join(ROOT, "scenes"))
    import gen_classroom_standin as gen
    from lumen_b200 import host
    out = os.environ.get("LUMEN_B200_SCENE_DIR") or os.path.join(ROOT, "scenes", "_generated", "classroom_standin")
    try:
        os.makedirs(out, exist_ok=True)
        path, _ = gen.generate(out)
    except OSError:
        path, _ = gen.generate(os.path.join(tempfile.gettempdir(), "lumen_b200_classroom_standin"))
    return host.Scene(path, width or WIDTH, height or HEIGHT)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.proc, self.path = None, tempfile.mktemp(suffix=".csv")
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def profile_file(name):
    """A committed summary of an ncu capture (profiles/<name>) + whether the CUDA sources changed since it was taken."""
    try:
        from srchash import csrc_sha
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        d["stale"] = d.get("csrc_sha") != csrc_sha()
        return d
    except Exception:
        return {}


def host_threads():
    return len(os.sched_getaffinity(0))  # every host thread this process may use (torchrun exports OMP_NUM_THREADS=1: not inherited)


class CpuReference:
    """The path on the host cores: the reference's own shaders compiled for the CPU (oracle/_ref/libglslref.so) when that library
    travelled with the snapshot, else the oracle port. Lumen itself needs a Vulkan RT GPU and has no CPU path."""

    def __init__(self, scene):
        from oracle import pyglslref as pr
        from oracle import pyoracle as po
        self.orc = po.OracleScene(scene)
        self.ref = pr.RefScene(scene, self.orc) if pr.available() else None
        self.kind = "reference" if self.ref else "port"
        self.what = ("the reference's path.rgen + includes, ray.rchit, ray.rmiss translated from the unmodified GLSL (oracle/glslref) and compiled; "
                     "ray/triangle intersection (the Vulkan driver's part) from the oracle's CPU LBVH") if self.ref else "oracle port (oracle/liboracle.so)"

    def frame(self, pc, ubo, frame, threads):
        """-> (rays, seconds) of one frame"""
        import numpy as np
        t0 = time.perf_counter()
        if self.ref:
            _, rays = self.ref.render(pc, ubo, frame, 1, rgba=np.zeros((pc.size_y, pc.size_x, 4), dtype=np.float32), threads=threads)
            n = int(rays.sum())
        else:
            _, st = self.orc.render_frame_raw(pc, ubo, frame, threads)
            n = st.rays
        return n, time.perf_counter() - t0


def run_reference(args, rank):
    """--impl reference: the reference's implementation of the path on the host cores (CpuReference), every host thread, one frame
    of the same 1920x1080 depth-8 workload per step."""
    if rank != 0:
        return
    scene = load_scene()
    cpu = CpuReference(scene)
    pc, ubo = scene.make_pc(MAX_DEPTH, True), scene.make_ubo()
    threads = host_threads()
    frame = 0
    for _ in range(args.warmup):
        cpu.frame(pc, ubo, frame, threads)
        frame += 1
    rays = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n, _ = cpu.frame(pc, ubo, frame, threads)
        rays += n
        frame += 1
    wall = time.perf_counter() - t0
    value = rays / wall / 1e6
    sample = f"{args.steps} steps x 1 frame at {WIDTH}x{HEIGHT} (the full workload, one sample per pixel per step), depth {MAX_DEPTH}; {cpu.what}"
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(w=WIDTH, h=HEIGHT), "max_depth": MAX_DEPTH, "width": WIDTH, "height": HEIGHT,
                   "triangles": int(scene.info.n_triangles), "frames_per_step": 1, "sample": sample},
        "spp_per_s": args.steps / wall,
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": cpu.kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def config5_roofline(integrator, local_rank, hbm_peak, peak_src):
    """BASELINE config 5 in this run: 10 M-triangle torus grid, closest hits of 2^24 incoherent rays (the BVH, 590 MB, does not fit
    the L2: the HBM-sized case). Algorithmic bytes counted by the kernel itself."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "scenes"))
    import gen_torus_grid as gen
    K, n = 10, 1 << 24
    scene = gen.make_scene(K, 100, 50, 64, 64)
    dev = integrator.Device(local_rank)
    try:
        dev.upload_scene(scene.desc)
        dev.build_accel()
        b = dev.stats()
        rays = torch.from_numpy(gen.random_rays(n, K)).cuda()
        hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        reps = 4
        dev.trace_closest_device(rays.data_ptr(), n, hits.data_ptr(), 1)  # warm-up
        dev.reset_stats()
        ms = dev.trace_closest_device(rays.data_ptr(), n, hits.data_ptr(), reps)
        s = dev.stats()
        first = hits.clone()
        # the same batch ordered by (origin cell, direction octant) INSIDE the timed launch (lmb_trace_closest_device_ex sort_rays = 1):
        # paid for itself while the walker's shared memory left the L1 28 KB (1483 -> 1912 Mrays/s); with the L1 at 92 KB it no longer does
        dev.trace_closest_device(rays.data_ptr(), n, hits.data_ptr(), 1, sort_rays=True)  # warm-up (allocates the sort scratch)
        ms_sorted = dev.trace_closest_device(rays.data_ptr(), n, hits.data_ptr(), reps, sort_rays=True)
        same = bool(torch.equal(first.view(torch.int32), hits.view(torch.int32)))
        traced = max(s.rays_closest, 1)
        bytes_per_ray = (s.nodes_visited * NODE_BYTES + s.tris_tested * TRI_BYTES) / traced + RAY_BYTES + HIT_BYTES
        mrays = n * reps / ms / 1e3
        achieved = bytes_per_ray * mrays / 1e3
        return {"bound": "hbm", "kernel": "k_trace_array (the same 8-wide walker over a ray array)", "workload": f"{int(scene.info.n_triangles)} triangles (torus grid {K}^3), {n} incoherent rays, closest hit",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "peak_source": peak_src, "mrays_per_s": mrays,
                "avg_launch_ms": ms / reps, "algorithmic_bytes_per_launch": bytes_per_ray * n, "bytes_per_ray": bytes_per_ray,
                "nodes_per_ray": s.nodes_visited / traced, "tris_per_ray": s.tris_tested / traced, "traffic": profile_file("ktrace_config5_traffic.json").get("dram_bytes_per_launch"),
                "traffic_stale": profile_file("ktrace_config5_traffic.json").get("stale"),
                "ray_order": "as given (uniformly random origins and directions)",
                "sorted": {"mrays_per_s": n * reps / ms_sorted / 1e3, "avg_launch_ms": ms_sorted / reps, "frac": bytes_per_ray * (n * reps / ms_sorted / 1e3) / 1e3 / hbm_peak,
                           "ray_order": "counting sort by (origin cell, direction octant), 15-bit keys, inside the timed launch (sort_rays = 1)"},
                "hits_identical_sorted_vs_unsorted": same,
                "lbvh_build_ms": {"total": b.ms_build_accel, "morton": b.ms_build_morton, "sort": b.ms_build_sort, "tree": b.ms_build_tree,
                                  "refit_pack": b.ms_build_refit, "wide": b.ms_build_wide}}
    finally:
        dev.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=16, help="frames (samples per pixel) one step renders in one wavefront batch per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bdpt", action="store_true", help="skip the BDPT leg (4 frames of lmb_render_bdpt on the same workload, N = 1 only)")
    ap.add_argument("--no-config5", action="store_true", help="skip the 10 M-triangle incoherent-ray leg (N = 1 only)")
    ap.add_argument("--no-config4", action="store_true", help="skip the 3840x2160 pixel x sample shard leg")
    ap.add_argument("--pixel-shards", type=int, default=1, help="P: ranks form a P x (N/P) grid of interleaved-row pixel shards x sample shards (headline arm)")
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    args = ap.parse_args()
    globals().update(WIDTH=args.width, HEIGHT=args.height)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from lumen_b200 import integrator, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator comes up; stdout carries exactly one JSON line,
        # so file descriptor 1 points at stderr until the communicators exist. torch.distributed carries only the harness here
        # (barriers, max over ranks, handing the communicator id round); the film exchange is the library's own (lmb_comm_*).
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        warm = torch.zeros(1, device="cuda")
        dist.all_reduce(warm)
        torch.cuda.synchronize()

    scene = load_scene()
    dev = integrator.Device(local_rank)
    dev.upload_scene(scene.desc)
    dev.build_accel()
    if world > 1:
        try:
            idt = torch.zeros(integrator.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(integrator.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            dev.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    fps = args.frames_per_step
    pc, ubo = scene.make_pc(MAX_DEPTH, True), scene.make_ubo()
    build = dev.stats()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_arm(width, height, frames_per_step, pixel_shards, steps, warmup):
        """K timed steps of the hot path at width x height; returns the device-resident and the host-buffer (e2e) measurement."""
        p_shard, s_shard, n_pshards, n_sshards = sharding.grid_of_rank(rank, world, pixel_shards)
        dev.set_pixel_shard(p_shard, n_pshards)
        dev.init(width, height, frames_per_step)
        a_pc, a_ubo = pc, ubo
        if (width, height) != (WIDTH, HEIGHT):
            sc2 = load_scene(width, height)
            a_pc, a_ubo = sc2.make_pc(MAX_DEPTH, True), sc2.make_ubo()
        pinned = torch.empty((height, width, 4), dtype=torch.float32, pin_memory=True)
        resident = torch.empty((height, width, 4), dtype=torch.float32, device="cuda") if world > 1 else None
        state = {"frame": 0}

        def step(download):
            first = state["frame"] + s_shard
            dev.render(a_pc, a_ubo, first, frames_per_step, n_sshards, integrator.FILM_SUM)  # frames first, first + S, ... of this rank's rows
            state["frame"] += frames_per_step * n_sshards
            if world > 1:
                # ONE fp32 sum-reduce to rank 0 (rgb sums + valid-sample counts) over NVLink + the "/ count" epilogue, on the context's comm
                # stream over a snapshot of the film; the film is cleared in stream order and the next step renders meanwhile. The job's
                # result is one image: rank 0 receives it (and, end to end, copies it home); the other ranks only contribute.
                dev.film_reduce(0, (pinned.data_ptr() if download else resident.data_ptr()) if rank == 0 else None, clear_film=True)
            else:
                dev.resolve()
                if download:
                    # the user-facing progressive read-back: film snapshot + copy-engine transfer into pinned memory, overlapped with
                    # the next step's rendering (lmb_download_async); timed() waits for the last transfer inside the timed region
                    dev.download_async(pinned.data_ptr())
                dev.clear_film()

        def timed(download):
            dev.clear_film()
            for _ in range(warmup):
                step(download)
            dev.sync()
            dev.reset_stats()
            flush.fill_(1)  # evict the BVH / film from L2 before the timed region; per-step state (>1 GB) streams anyway
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                step(download)
            dev.sync()  # every queued kernel, reduce and film transfer
            barrier()
            dt = time.perf_counter() - t0
            st = dev.stats()
            t = torch.tensor([dt, float(st.rays), float(st.kernel_launches), float(st.ms_render)], dtype=torch.float64, device="cuda")
            if world > 1:
                tmax, tsum = t.clone(), t.clone()
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
                return dict(dt=tmax[0].item(), rays=tsum[1].item(), launches=tsum[2].item(), ms_render=tmax[3].item(), st=st)
            return dict(dt=dt, rays=float(st.rays), launches=float(st.kernel_launches), ms_render=float(st.ms_render), st=st)

        clocks = ClockSampler(local_rank) if rank == 0 else None
        res = timed(download=False)
        clk = clocks.stop() if clocks else None
        e2e = timed(download=True)
        return dict(res=res, e2e=e2e, clk=clk, n_pshards=n_pshards, n_sshards=n_sshards, s_shard=s_shard, pinned=pinned, resident=resident,
                    pc=a_pc, ubo=a_ubo)

    arm = run_arm(WIDTH, HEIGHT, fps, args.pixel_shards, args.steps, args.warmup)
    res, e2e, clk, n_pshards, n_sshards = arm["res"], arm["e2e"], arm["clk"], arm["n_pshards"], arm["n_sshards"]

    # ---- N-rank film == the same frames on rank 0 alone, outside the timed regions
    film_check = None
    if world > 1 and n_pshards == 1:
        first = 500_000
        dev.clear_film()
        dev.render(pc, ubo, first + arm["s_shard"], fps, n_sshards, integrator.FILM_SUM)
        dev.film_reduce(0, arm["pinned"].data_ptr() if rank == 0 else None, clear_film=True)
        dev.sync()
        barrier()
        if rank == 0:
            reduced = arm["pinned"].numpy().copy()
            dev.render(pc, ubo, first, fps * n_sshards, 1, integrator.FILM_SUM)
            dev.resolve()
            alone = dev.download()
            dev.clear_film()
            rel = np.abs(reduced[..., :3] - alone[..., :3]) / np.maximum(np.abs(alone[..., :3]), 1e-3)
            film_check = {"max_rel_diff": float(rel.max()), "pixels_within_1e-4": float((rel.max(axis=2) <= 1e-4).mean()), "frames": fps * n_sshards,
                          "alpha_is_one": bool((reduced[..., 3] == 1.0).all()),
                          "note": "resolved film of the N-rank reduce (lmb_film_reduce, root 0) vs the same frames rendered by rank 0 alone (sum film + resolve); the fp32 sums differ only in association"}
        barrier()

    # ---- roofline of the dominant kernel (traversal), from a separately profiled pass of this run: per-stage CUDA events + the
    # kernel's own counters
    dev.set_profile_stages(True)
    dev.reset_stats()
    dev.clear_film()
    dev.render(pc, ubo, 1_000_000, fps, 1, integrator.FILM_SUM)
    ps = dev.stats()
    dev.set_profile_stages(False)
    dev.clear_film()
    n_sms = torch.cuda.get_device_properties(local_rank).multi_processor_count

    # ---- config 4's shape: 3840x2160, 2 pixel shards x N/2 sample shards, ONE frame per rank and step (every rank takes part)
    config4 = None
    if not args.no_config4:
        c4 = run_arm(3840, 2160, 1, 2 if world >= 2 else 1, 12, 3)
        r4, e4 = c4["res"], c4["e2e"]
        frames4 = 12 * c4["n_sshards"]
        config4 = {"value": r4["rays"] / r4["dt"] / 1e6, "unit": "Mrays/s", "ms_per_step": r4["dt"] / 12 * 1e3, "spp_per_s": frames4 / r4["dt"],
                   "e2e": {"value": e4["rays"] / e4["dt"] / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 52 + 492, "d2h_bytes_per_step": 3840 * 2160 * 16,
                           "ms_per_step": e4["dt"] / 12 * 1e3},
                   "device_ms_render_per_step": r4["ms_render"] / 12, "steps": 12, "warmup": 3, "n_gpus": world,
                   "workload": f"classroom-standin 3840x2160 depth 8, {c4['n_pshards']} pixel shard(s) (interleaved rows) x {c4['n_sshards']} sample shard(s), 1 frame per rank and step"
                               + (", films summed on rank 0 by one 133 MB lmb_film_reduce per step (overlapped with the next step)" if world > 1 else ""),
                   "note": "BASELINE config 4 (bedroom 4K, tile + sample sharding) on the stand-in geometry; scaling efficiency = value(N) / (N x value(1)) over the SCALE records"}
        dev.set_pixel_shard(0, 1)
        dev.init(WIDTH, HEIGHT, fps)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        trav_ms = ps.ms_extend
        trav_bytes = ps.nodes_visited * NODE_BYTES + ps.tris_tested * TRI_BYTES + ps.rays * (RAY_BYTES + HIT_BYTES)
        n_trav_launches = MAX_DEPTH
        cal = profile_file("ktrace_calibration.json")
        traffic = profile_file("ktrace_dram_traffic.json")
        sm_mhz = (clk or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
        issue_peak = n_sms * 4 * sm_mhz * 1e6 / 1e9  # G warp-instructions / s: 4 schedulers per SM, one warp instruction per cycle each
        w = cal.get("warp_inst_per", {})
        warp_inst = (ps.trace_warp_iters * w.get("iters", 0.0) + ps.rays * w.get("rays", 0.0)) if w else None
        issue_achieved = warp_inst / (trav_ms * 1e-3) / 1e9 if (warp_inst and trav_ms > 0) else None
        mem_achieved = trav_bytes / (trav_ms * 1e-3) / 1e9 if trav_ms > 0 else 0.0
        roofline = {
            "bound": "issue", "kernel": "k_trace (persistent traversal of the 8-wide BVH: continuation + shadow + MIS-probe rays)",
            "achieved": issue_achieved, "peak": issue_peak, "unit": "Gwarp-inst/s", "frac": (issue_achieved / issue_peak) if issue_achieved else None,
            "peak_source": f"{n_sms} SMs x 4 schedulers x {sm_mhz:.0f} MHz (SM clock sampled by nvidia-smi during the timed region of this run)",
            "warp_inst_per_ray": (warp_inst / max(ps.rays, 1)) if warp_inst else None,
            "counters": {"loop_trips": int(ps.trace_warp_iters), "rays": int(ps.rays)},
            "calibration": {"warp_inst_per": w, "max_rel_residual": cal.get("max_rel_residual"), "source": "profiles/ktrace_calibration.json (ncu smsp__inst_executed.sum per launch fitted on the kernel's counters)",
                            "stale": cal.get("stale", True)},
            "avg_launch_ms": trav_ms / n_trav_launches,
            "traffic": traffic.get("dram_bytes_per_launch"), "traffic_stale": traffic.get("stale", True),
            # secondary: the memory view. The wide BVH (~18 MB) is L2-resident by design, so these bytes are served by L1 / L2, not by HBM
            "algorithmic": {"achieved": mem_achieved, "peak": hbm_peak, "unit": "GB/s", "frac_of_hbm_peak": mem_achieved / hbm_peak, "peak_source": peak_src,
                            "bytes_per_launch": trav_bytes / n_trav_launches, "bytes_per_ray": trav_bytes / max(ps.rays, 1),
                            "nodes_per_ray": ps.nodes_visited / max(ps.rays, 1), "tris_per_ray": ps.tris_tested / max(ps.rays, 1)},
            "ncu": {k: traffic.get(k) for k in ("issue_active_pct", "alu_pipe_pct", "fma_pipe_pct", "active_lanes_per_instruction", "source")},
            "stage_ms": {"trace": ps.ms_extend, "shade": ps.ms_shade, "connect": ps.ms_connect, "raygen_sky_film": ps.ms_film, "total": ps.ms_render},
            "note": "issue bound: warp instructions = the kernel's own loop-trip counter and ray count of THIS run x the per-trip / per-ray instruction costs fitted on the committed ncu launch list; "
                    "algorithmic bytes = wide nodes visited*80 + triangles tested*48 + rays*48 (DESIGN.md), served by L1/L2 here (traffic = ncu dram bytes per launch)",
        }
        roofline5 = None
        if world == 1 and not args.no_config5:
            try:
                roofline5 = config5_roofline(integrator, local_rank, hbm_peak, peak_src)
            except Exception as e:  # the headline line must not depend on this leg
                roofline5 = {"error": str(e)}
        cpu_baseline = None
        if not args.no_cpu_baseline:
            cpu = CpuReference(scene)
            threads = host_threads()
            cpu.frame(pc, ubo, 0, threads)  # warm-up (page-in, thread pool)
            n, c_rays, c_secs = 0, 0, 0.0
            while c_secs < 12.0 and n < 400:  # bounded sample: about 12 s of CPU work on all host threads
                r_, s_ = cpu.frame(pc, ubo, 1 + n, threads)
                c_rays += r_
                c_secs += s_
                n += 1
            cpu_baseline = {"value": c_rays / c_secs / 1e6, "unit": "Mrays/s", "cores": threads, "kind": cpu.kind,
                            "sample": f"{n} frames at {WIDTH}x{HEIGHT} (the full workload, one sample per pixel each), depth {MAX_DEPTH}, {c_secs:.1f} s; {cpu.what}"}
        # ---- the sibling integrator of SURVEY.md 8f rank 3 on the same workload (reported beside the headline, never part of it)
        bdpt = None
        if world == 1 and not args.no_bdpt:
            try:
                from lumen_b200._ctypes_types import PCBdpt
                pcb = PCBdpt.from_path_pc(pc)
                dev.clear_film()
                dev.render_bdpt(pcb, ubo, 0, 2)  # warm-up: allocates the vertex / slot buffers
                dev.reset_stats()
                dev.render_bdpt(pcb, ubo, 2, 4)
                bs = dev.stats()
                brays = bs.rays_closest + bs.rays_shadow
                bdpt = {"value": brays / bs.ms_render / 1e3, "unit": "Mrays/s", "spp_per_s": 4 / (bs.ms_render * 1e-3), "ms_per_frame": bs.ms_render / 4,
                        "rays_per_pixel": brays / 4 / (WIDTH * HEIGHT), "frames": 4, "gpu_launches": int(bs.kernel_launches),
                        "note": "lmb_render_bdpt (bdpt.rgen + bdpt_commons.glsl restated, DESIGN.md section 8), same scene / size / max_depth, device time (CUDA events)"}
            except Exception as e:  # the headline line must not depend on this leg
                bdpt = {"error": str(e)}
        dt, rays = res["dt"], res["rays"]
        total_frames = args.steps * fps * n_sshards  # whole-image frames: a pixel shard renders 1/P of each
        so_exists = lambda p: os.path.exists(os.path.join(ROOT, p))  # noqa: E731
        line = {
            "metric": "Mrays/s", "value": rays / dt / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak" if n_pshards == 1 else "mixed (pixel shards split the image, sample shards add frames)", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(w=WIDTH, h=HEIGHT), "max_depth": MAX_DEPTH, "frames_per_step_per_gpu": fps, "triangles": int(scene.info.n_triangles),
                       "sharding": (f"{n_pshards} pixel shard(s) (interleaved rows) x {n_sshards} sample shard(s) (frame index mod {n_sshards}), full scene + BVH replica per GPU, one lmb_film_reduce to rank 0 (ncclReduce fp32 + resolve epilogue, overlapped with the next step) per step"
                                    if world > 1 else "single GPU"), "width": WIDTH, "height": HEIGHT,
                       "l2": f"256 MB flush before the timed region; per-step wavefront state (~{SLOT_BYTES * fps * WIDTH * HEIGHT / 1e9:.1f} GB) exceeds the 126 MB L2, the ~18 MB BVH stays L2-resident by design",
                       "libraries": {"liblumen_b200.so": "built in-tree (nvcc sm_100a)", "liblumen_host.so / liboracle.so / libglslref.so": "prebuilt in the build container (they compile against the reference checkout's third-party parsers / shader sources, which the GPU box does not have) and shipped with the snapshot",
                                     "libglslref_present": so_exists("oracle/_ref/libglslref.so")}},
            "spp_per_s": total_frames / dt,
            "rays_per_path": rays / (total_frames * WIDTH * HEIGHT),
            "device_ms_render_per_step": res["ms_render"] / args.steps,
            "e2e": {"value": e2e["rays"] / e2e["dt"] / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 52 + 492, "d2h_bytes_per_step": WIDTH * HEIGHT * 16,
                    "ms_per_step": e2e["dt"] / args.steps * 1e3},
            "gpu_launches": int(res["launches"]),
            "clocks": clk,
            "roofline": roofline,
            "roofline_config5": roofline5,
            "config4": config4,
            "film_check": film_check,
            "cpu_baseline": cpu_baseline,
            "bdpt": bdpt,
            "lbvh_build_ms": {"total": build.ms_build_accel, "morton": build.ms_build_morton, "sort": build.ms_build_sort, "tree": build.ms_build_tree,
                              "refit_pack": build.ms_build_refit},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dev.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
