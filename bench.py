#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 path integrator (BASELINE.json metric: Mrays/s and spp/s, classroom 1080p).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): the classroom configuration of BASELINE.json -- 1920x1080, max depth 8, sun + sky, Disney /
diffuse / conductor materials. The reference's classroom meshes are not in its tree (scenes/classroom/how-to-obtain.txt),
so the geometry is the labelled procedural stand-in of scenes/gen_classroom_standin.py, loaded through the same
Mitsuba-XML path. One STEP = one pass of the hot path over one batch: every rank renders `frames_per_step` frames
(samples per pixel) of the full image into a cleared film, the per-GPU sums are combined (NCCL all-reduce when N > 1) and
resolved. Ranks take disjoint frame indices (frame = first + rank + k*N): per-GPU work is fixed -> weak scaling.

`value`  = traced rays of all ranks / time of K steps, scene + film resident in HBM, timed between barriers +
           torch.cuda.synchronize(), max over ranks.
`e2e`    = same metric through the public C-ABI calls with HOST buffers inside the timed region: per step the push
           constants + camera UBO are handed over from host memory (lmb_render copies them) and the resolved RGBA32F film is
           downloaded to pinned host memory (lmb_download).
`roofline` is for the dominant kernel (k_trace, BVH traversal of all ray types), from a separately profiled pass.
`cpu_baseline` = the CPU oracle (C++/glm restatement of the reference shaders, OpenMP) on a bounded sample.
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

WIDTH, HEIGHT, MAX_DEPTH = 1920, 1080, 8
WORKLOAD = "classroom-standin {w}x{h} depth 8 (procedural stand-in for scenes/classroom, whose meshes are not in the reference tree)"
NODE_BYTES, TRI_BYTES, RAY_BYTES, HIT_BYTES = 80, 48, 32, 16  # 8-wide compressed node, 3 x float4 triangle, ray in, hit out


def load_scene():
    sys.path.insert(0, os.path.join(ROOT, "scenes"))
    import gen_classroom_standin as gen
    from lumen_b200 import host
    out = os.environ.get("LUMEN_B200_SCENE_DIR") or os.path.join(ROOT, "scenes", "_generated", "classroom_standin")
    try:
        os.makedirs(out, exist_ok=True)
        path, _ = gen.generate(out)
    except OSError:
        path, _ = gen.generate(os.path.join(tempfile.gettempdir(), "lumen_b200_classroom_standin"))
    return host.Scene(path, WIDTH, HEIGHT)


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


def ncu_capture():
    """The committed ncu --set full capture of k_trace (profiles/ktrace_dram_traffic.json, written by tools/ncu_summary.py
    from the capture of tools/gpu_round.sh): dram bytes per launch and the issue / pipe utilisation of the same launches."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ktrace_dram_traffic.json")))
    except Exception:
        return {}


def ncu_traffic():
    v = ncu_capture().get("dram_bytes_per_launch")
    return float(v) if v is not None else None


def run_reference(args, rank):
    """--impl reference: the reference's algorithm on the host cores. Lumen itself needs a Vulkan RT GPU and has no CPU
    path, so this arm times the oracle port (oracle/liboracle.so) with every host thread on a bounded sample per step:
    one frame of the same 1920x1080 depth-8 workload."""
    if rank != 0:
        return
    from oracle import pyoracle as po
    scene = load_scene()
    orc = po.OracleScene(scene)
    pc, ubo = scene.make_pc(MAX_DEPTH, True), scene.make_ubo()
    threads = len(os.sched_getaffinity(0))  # every host thread this process may use (torchrun exports OMP_NUM_THREADS=1: do not inherit it)
    # bounded sample: a strip of scanlines would bias the ray mix, so the sample is the full view at half resolution
    pc.size_x, pc.size_y = WIDTH // 2, HEIGHT // 2
    frame = 0
    for _ in range(args.warmup):
        orc.render_frame_raw(pc, ubo, frame, threads)
        frame += 1
    rays, secs = 0, 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, st = orc.render_frame_raw(pc, ubo, frame, threads)
        rays += st.rays
        secs += st.seconds
        frame += 1
    wall = time.perf_counter() - t0
    value = rays / wall / 1e6
    sample = f"{args.steps} steps x 1 frame at {pc.size_x}x{pc.size_y} (full view, half resolution), depth {MAX_DEPTH}"
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(w=WIDTH, h=HEIGHT), "max_depth": MAX_DEPTH, "sample": sample},
        "spp_per_s": args.steps / wall * (pc.size_x * pc.size_y) / (WIDTH * HEIGHT),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=16, help="frames (samples per pixel) one step renders in one wavefront batch per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bdpt", action="store_true", help="skip the BDPT leg (4 frames of lmb_render_bdpt on the same workload, N = 1 only)")
    ap.add_argument("--pixel-shards", type=int, default=1, help="P: ranks form a P x (N/P) grid of interleaved-row pixel shards x sample shards (config 4)")
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
    from lumen_b200 import integrator

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator comes up; stdout carries exactly one JSON line,
        # so file descriptor 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    scene = load_scene()
    dev = integrator.Device(local_rank)
    dev.upload_scene(scene.desc)
    dev.build_accel()
    fps = args.frames_per_step
    from lumen_b200 import sharding
    p_shard, s_shard, n_pshards, n_sshards = sharding.grid_of_rank(rank, world, args.pixel_shards)
    dev.set_pixel_shard(p_shard, n_pshards)
    dev.init(WIDTH, HEIGHT, fps)
    pc, ubo = scene.make_pc(MAX_DEPTH, True), scene.make_ubo()
    build = dev.stats()

    # zero-copy torch view of the film for the NCCL all-reduce
    ptr, n_floats = dev.film_device_ptr()

    class _Film:
        __cuda_array_interface__ = {"shape": (HEIGHT, WIDTH, 4), "typestr": "<f4", "data": (ptr, False), "version": 2}

    film_t = torch.as_tensor(_Film(), device=torch.device("cuda", local_rank))
    pinned = torch.empty((HEIGHT, WIDTH, 4), dtype=torch.float32, pin_memory=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = {"frame": 0}

    def step(download):
        dev.clear_film()
        first = state["frame"] + s_shard
        dev.render(pc, ubo, first, fps, n_sshards, integrator.FILM_SUM)  # frames first, first + S, ... of this rank's rows
        state["frame"] += fps * n_sshards
        if world > 1:
            dist.all_reduce(film_t)  # fp32 sum of rgb and of the per-pixel valid-sample count over NVLink
            torch.cuda.synchronize()
        dev.resolve()
        if download:
            # the user-facing progressive read-back: film snapshot + copy-engine transfer into pinned memory, overlapped with
            # the next step's rendering (lmb_download_async); timed() waits for the last transfer inside the timed region
            dev.download_async(pinned.data_ptr())

    def timed(download):
        for _ in range(args.warmup):
            step(download)
        dev.reset_stats()
        flush.fill_(1)  # evict the BVH / film from L2 before the timed region; per-step state (>1 GB) streams anyway
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(download)
        dev.sync()  # every queued kernel and film transfer
        barrier()
        dt = time.perf_counter() - t0
        st = dev.stats()
        t = torch.tensor([dt, float(st.rays), float(st.kernel_launches), float(st.ms_render)], dtype=torch.float64, device="cuda")
        if world > 1:
            tmax, tsum = t.clone(), t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            return tmax[0].item(), tsum[1].item(), tsum[2].item(), tmax[3].item(), st
        return dt, float(st.rays), float(st.kernel_launches), float(st.ms_render), st

    clocks = ClockSampler(local_rank) if rank == 0 else None
    dt, rays, launches, ms_render, st = timed(download=False)
    clk = clocks.stop() if clocks else None
    dt_e2e, rays_e2e, _, _, _ = timed(download=True)

    # ---- roofline of the dominant kernels (traversal), from a separately profiled pass: per-stage CUDA events
    dev.set_profile_stages(True)
    dev.reset_stats()
    dev.clear_film()
    dev.render(pc, ubo, 1_000_000, fps, 1, integrator.FILM_SUM)
    ps = dev.stats()
    dev.set_profile_stages(False)

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
        achieved = trav_bytes / (trav_ms * 1e-3) / 1e9 if trav_ms > 0 else 0.0
        roofline = {
            "bound": "hbm", "kernel": "k_trace (persistent traversal of the 8-wide BVH: continuation + shadow + MIS-probe rays)", "achieved": achieved,
            "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": ncu_traffic(), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": trav_bytes / n_trav_launches, "avg_launch_ms": trav_ms / n_trav_launches,
            "bytes_per_ray": trav_bytes / max(ps.rays, 1), "nodes_per_ray": ps.nodes_visited / max(ps.rays, 1), "tris_per_ray": ps.tris_tested / max(ps.rays, 1),
            # second ceiling of SURVEY.md 8d, from the committed ncu capture (not re-measured in this run): instruction issue
            "issue": {k: ncu_capture().get(k) for k in ("issue_active_pct", "alu_pipe_pct", "fma_pipe_pct", "active_lanes_per_instruction", "source")},
            "stage_ms": {"trace": ps.ms_extend, "shade": ps.ms_shade, "connect": ps.ms_connect, "raygen_sky_film": ps.ms_film, "total": ps.ms_render},
            "note": "algorithmic bytes = wide nodes visited*80 + triangles tested*48 + rays*48 (DESIGN.md); the ~18 MB wide BVH is L2-resident by design, so these bytes are served by L1/L2 and the kernel is issue bound, not DRAM bound: see traffic (ncu dram bytes per launch) against algorithmic_bytes_per_launch",
        }
        cpu_baseline = None
        if not args.no_cpu_baseline:
            from oracle import pyoracle as po
            orc = po.OracleScene(scene)
            pcs = scene.make_pc(MAX_DEPTH, True)
            pcs.size_x, pcs.size_y = WIDTH // 2, HEIGHT // 2
            host_threads = len(os.sched_getaffinity(0))
            orc.render_frame_raw(pcs, ubo, 0, host_threads)  # warm-up (page-in, thread pool)
            _, cst = orc.render_frame_raw(pcs, ubo, 1, host_threads)
            n = 1
            while cst.seconds < 12.0 and n < 400:  # bounded sample: about 12 s of CPU work on all host threads
                _, c2 = orc.render_frame_raw(pcs, ubo, 1 + n, host_threads)
                cst.rays_closest += c2.rays_closest
                cst.rays_shadow += c2.rays_shadow
                cst.rays_probe += c2.rays_probe
                cst.seconds += c2.seconds
                n += 1
            cpu_baseline = {"value": cst.rays / cst.seconds / 1e6, "unit": "Mrays/s", "cores": cst.threads, "kind": "port",
                            "sample": f"{n} frames at {pcs.size_x}x{pcs.size_y} (full view, half resolution), depth {MAX_DEPTH}, {cst.seconds:.1f} s"}
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
        total_frames = args.steps * fps * n_sshards  # whole-image frames: a pixel shard renders 1/P of each
        line = {
            "metric": "Mrays/s", "value": rays / dt / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak" if n_pshards == 1 else "mixed (pixel shards split the image, sample shards add frames)", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(w=WIDTH, h=HEIGHT), "max_depth": MAX_DEPTH, "frames_per_step_per_gpu": fps, "triangles": int(scene.info.n_triangles),
                       "sharding": (f"{n_pshards} pixel shard(s) (interleaved rows) x {n_sshards} sample shard(s) (frame index mod {n_sshards}), full scene + BVH replica per GPU, fp32 film all-reduce per step"
                                    if world > 1 else "single GPU"), "width": WIDTH, "height": HEIGHT,
                       "l2": f"256 MB flush before the timed region; per-step wavefront state (~{0.33 * fps * WIDTH * HEIGHT / 1e6 / 1e3:.1f} GB) exceeds the 126 MB L2, the 15 MB BVH stays L2-resident by design"},
            "spp_per_s": total_frames / dt,
            "rays_per_path": rays / (total_frames * WIDTH * HEIGHT),
            "device_ms_render_per_step": ms_render / args.steps,
            "e2e": {"value": rays_e2e / dt_e2e / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 52 + 492, "d2h_bytes_per_step": WIDTH * HEIGHT * 16,
                    "ms_per_step": dt_e2e / args.steps * 1e3},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "bdpt": bdpt,
            "lbvh_build_ms": {"total": build.ms_build_accel, "morton": build.ms_build_morton, "sort": build.ms_build_sort, "tree": build.ms_build_tree,
                              "refit_pack": build.ms_build_refit},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
