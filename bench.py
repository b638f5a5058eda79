#!/usr/bin/env python3
"""bench.py -- primary Mrays/s of the level-set ray tracer on the BASELINE.json configuration.

Workload ("c2"): nanovdb createLevelSetTorus(R=650, r=325, voxel 1, half-width 3) ~ 50 M active voxels (0.58 GB NanoVDB
grid, built on the GPU by the library's own builder), 1920x1080, 1 spp, DiffuseShader, vdb_render's perspective camera
at (0, 1.5R, 3(R+r)) looking at the origin.

  value    whole-frame primary rays / device time, grid and film resident in HBM (CUDA events, max over ranks)
  e2e      the same metric through the C ABI with a pinned HOST film: H2D of the film + kernel + D2H inside the timing
  roofline algorithmic bytes per ray (counted by an instrumented launch, SURVEY 8d formula) x rays / kernel time
  cpu_baseline  the reference's own CPU path (oracle/_ref, all host threads) on a bounded sample of the same workload

N > 1 (torchrun, one rank per GPU): the grid is replicated, the film is split into 64x60-pixel tiles interleaved over
the ranks (vdbrt_partition), and every step ends with an NCCL gather of the owned tiles to rank 0 ("strong" scaling:
the frame is fixed).  `--impl reference` times the reference CPU implementation alone (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (major, minor, width, height)  -- level-set torus
    "c2": (650.0, 325.0, 1920, 1080),
    "c2-small": (160.0, 80.0, 640, 360),     # quick functional check, not a bench line
    # BASELINE config 4: union of 10 000 random spheres (~1 B active voxels, ~12 GB grid replicated per GPU), 3840x2160;
    # (n spheres, extent) in place of the radii; camera at (0, 0, 3*2048)
    "c4": (10000, 1988.0, 3840, 2160),
    "c4-small": (300, 600.0, 1280, 720),
}
TILE_W, TILE_H = 64, 60


def camera_args(R, r, workload="c2"):
    if workload.startswith("c4"):
        return (0.0, 0.0, 3.0 * (r + 60.0)), (0.0, 0.0, 0.0)     # r = extent of the sphere centres; 3*2048 for the full set
    return (0.0, 1.5 * R, 3.0 * (R + r)), (0.0, 0.0, 0.0)


def build_grid(ctx, api, workload):
    R, r, _, _ = WORKLOADS[workload]
    if workload.startswith("c4"):
        return ctx.build_spheres(api.random_spheres(int(R), 20240607, r, 10.0, 60.0))
    return ctx.build_torus(R, r)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bytes_per_ray(c):
    """B_LS = 32 n_R + 16 n_U + 16 n_L + 12 n_V + 32 n_S + 16"""
    n = float(c["rays"])
    return (32 * c["root_probes"] + 16 * c["upper_probes"] + 16 * c["lower_probes"] + 12 * c["voxel_probes"]
            + 32 * c["stencil_refills"]) / n + 16.0


def cpu_reference(R, r, W, H, steps, warmup, sample_div=1, gpu_film=None):
    """the reference's own CPU implementation (oracle/_ref: unmodified OpenVDB LevelSetRayTracer, threaded) on a bounded
    sample of the workload: the same grid and camera at (W/div) x (H/div) pixels.  Falls back to the oracle port."""
    from tests import refapi
    from openvdb_b200 import _abi as abi
    w, h = max(W // sample_div, 1), max(H // sample_div, 1)
    tr, look = camera_args(R, r)
    cores = os.cpu_count() or 1
    if os.path.exists(refapi.REF_SO):
        ref = refapi.Ref()
        ref.set_threads(cores)
        g = ref.torus(R, r)
        d = refapi.camera_desc(w, h, translation=tr, lookat=look)
        sh = refapi.shader(abi.SHADER_DIFFUSE)
        film = refapi.new_film(w, h)
        times = []
        for it in range(warmup + steps):
            film[...] = (0, 0, 0, 1)
            t = ref.render_levelset(g, d, sh, film, threaded=True)
            if it >= warmup:
                times.append(t)
        kind = "reference"
        hits = int((film[..., :3].sum(axis=2) > 0).sum())
    else:
        from openvdb_b200 import api
        oracle = refapi.Oracle()
        ctx = api.Context(0)
        grid = ctx.build_torus(R, r)
        og = oracle.open(grid.download())
        cam = api.vdb_render_camera(w, h, tr, look)
        sh = api.make_shader(abi.SHADER_DIFFUSE)
        film = refapi.new_film(w, h)
        times = []
        for it in range(warmup + steps):
            film[...] = (0, 0, 0, 1)
            t0 = time.perf_counter()
            oracle.render_levelset(og, cam, sh, film, threads=cores)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        kind = "port"
        hits = int((film[..., :3].sum(axis=2) > 0).sum())
    t = float(np.median(times))
    parity = None
    if gpu_film is not None and gpu_film.shape == film.shape:
        # the checker's frame against the GPU's frame of the same workload (both start from a (0,0,0,1) film)
        bad = int((gpu_film != film).any(axis=2).sum())
        parity = {"against": kind, "pixels": int(w * h), "mismatched_pixels": bad, "max_abs_diff": float(np.abs(gpu_film - film).max())}
    return {"value": w * h / t / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind, "parity": parity,
            "sample": "%dx%d pixels of the same torus/camera (1/%d of the frame's rays), median of %d runs after %d warm-up, %d hit pixels"
                      % (w, h, sample_div * sample_div, steps, warmup, hits),
            "ms_per_sample": t * 1e3}


def extras(ctx, api, abi, torch):
    """informational timings of the other BASELINE configs (device-resident film, CUDA-event kernel time, 3 frames each):
    C3 fog sphere (1024^3 bbox, step 0.5, 1920x1080), C4 union of 10 000 spheres at 3840x2160 ('ms per 4K frame') and one
    GPU's share of C5 (level set + fog overlay, 16 samples per pixel)"""
    out = {}
    ls = ctx.build_sphere(509.0)
    fog = ctx.build_fog(ls)
    ls.free()
    W, H = 1920, 1080
    cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3 * 509.0), (0.0, 0.0, 0.0))
    vo = api.vol_opts_default()
    vo.primary_step = 0.5
    film = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    ms = []
    for _ in range(4):
        ctx.render_volume(fog, cam, vo, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE)
        ms.append(ctx.last_kernel_ms()[0])
    out["c3_fog_1080p"] = {"ms_per_frame": float(np.median(ms[1:])), "Mrays_per_s": W * H / float(np.median(ms[1:])) / 1e3,
                           "grid_gb": fog.info.bytes / 1e9, "alpha_sum": float(film[..., 3].sum().item())}
    fog.free()
    g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0))
    W, H = 3840, 2160
    cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
    film = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    opts = ctx.ls_opts(uniform_bg=True)
    ms = []
    for _ in range(4):
        ctx.render_levelset(g, cam, api.make_shader(abi.SHADER_DIFFUSE), film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=opts)
        ms.append(ctx.last_kernel_ms()[0])
    out["c4_levelset_4k"] = {"ms_per_frame": float(np.median(ms[1:])), "Mrays_per_s": W * H / float(np.median(ms[1:])) / 1e3,
                             "grid_gb": g.info.bytes / 1e9, "active_voxels": int(g.info.active_voxels),
                             "hit_pixels": int((film[..., :3].sum(dim=2) > 0).sum().item())}
    # C5 (extension): the same union under its own fog volume, 16 jittered samples per pixel, fog.over(level set).
    # The config names the whole 8-GPU box: this one GPU renders rank 0's share of an 8-way tile split, once.
    try:
        fog = ctx.build_fog(g)
        spp, share = 16, 8
        part = api.partition(0, share, TILE_W, TILE_H)
        f2 = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        ctx.render_levelset(g, cam, api.make_shader(abi.SHADER_DIFFUSE), film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE,
                            opts=ctx.ls_opts(spp=spp, seed=0, uniform_bg=True, part=part))
        ms_ls = ctx.last_kernel_ms()[0]
        vo = api.vol_opts_default(spp=spp, seed=0)
        vo.primary_step = 0.5
        vo.part = part
        ctx.render_volume(fog, cam, vo, f2.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE)
        ms_fog = ctx.last_kernel_ms()[0]
        ctx.film_over(f2.data_ptr(), film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE)
        ms_over = ctx.last_kernel_ms()[0]
        rays = W * H * spp // share
        out["c5_overlay_4k_16spp_rank0_of_8"] = {"ms_level_set": ms_ls, "ms_fog": ms_fog, "ms_over_whole_film": ms_over,
                                                 "ms_per_frame": ms_ls + ms_fog + ms_over, "primary_rays": 2 * rays,
                                                 "Mrays_per_s": 2 * rays / (ms_ls + ms_fog + ms_over) / 1e3, "fog_grid_gb": fog.info.bytes / 1e9}
        fog.free()
    except Exception as e:
        out["c5_overlay_4k_16spp_rank0_of_8"] = {"error": str(e)}
    g.free()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    R, r, W, H = WORKLOADS[args.workload]
    cb = cpu_reference(R, r, W, H, max(args.steps, 1), max(min(args.warmup, 2), 1))
    line = {"impl": "reference", "metric": "primary Mrays/s, level-set ray tracer", "value": cb["value"], "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_sample"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload + ": level-set torus R=%g r=%g, %dx%d, 1 spp, diffuse (bounded sample)" % (R, r, W, H)},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vdbrt", choices=["vdbrt", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational fog (C3) and 4K (C4) timings")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU frame assembly: render kernels store their tiles straight into rank 0's film over NVLink "
                         "(CUDA IPC mapping), or pack + NCCL gather + unpack")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from openvdb_b200 import api, _abi as abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libvdbrt.so has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warmup = max(args.warmup, 3)
    R, r, W, H = WORKLOADS[args.workload]
    tr, look = camera_args(R, r, args.workload)

    ctx = api.Context(local)
    # one non-default torch stream carries the library's kernels, torch's copies, NCCL and the timing events
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    grid = build_grid(ctx, api, args.workload)         # replicated on every GPU
    build_s = time.perf_counter() - t0
    cam = api.vdb_render_camera(W, H, tr, look)
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    part = api.partition(rank, world, TILE_W, TILE_H) if world > 1 else None
    film = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    bg = (0.0, 0.0, 0.0, 1.0)                           # a fresh tools::Film (RayTracer.h:235)
    opts = ctx.ls_opts(part=part, uniform_bg=True)
    opts.flags |= abi.ASYNC

    # multi-GPU frame assembly.  "peer": every rank's render kernel stores the tiles it owns straight into rank 0's film (CUDA IPC
    # mapping, NVLink peer stores) and a 4-byte all-reduce orders the streams -- compute and gather are one kernel.  "nccl":
    # openvdb_b200/frame.py packs the owned tiles, gathers them with one NCCL collective and unpacks on rank 0.
    from openvdb_b200.frame import TileGather
    peer = world > 1 and args.gather == "peer"
    gather = TileGather(H, W, TILE_H, TILE_W, rank, world, "cuda") if (world > 1 and not peer) else None
    token = torch.zeros(1, dtype=torch.float32, device="cuda")
    film_ptr = film.data_ptr()
    shared = None
    if peer:
        def exchange(h):
            t = torch.from_numpy(h.copy()).cuda()
            dist.broadcast(t, src=0)
            return t.cpu().numpy()
        shared = api.SharedFilm(ctx, H, W, rank, exchange)
        film_ptr = shared.ptr

    def step():
        ctx.render_levelset(grid, cam, sh, film_ptr, width=W, height=H, memspace=abi.MEM_DEVICE, bg=bg, opts=opts)
        if gather:
            gather.gather(film)
        elif peer:
            dist.all_reduce(token)          # stream-ordered: rank 0 continues only after every rank's kernel has stored its tiles

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    for k in range(args.steps):
        ev[k][0].record(stream)
        ctx.render_levelset(grid, cam, sh, film_ptr, width=W, height=H, memspace=abi.MEM_DEVICE, bg=bg, opts=opts)
        ev[k][1].record(stream)
        if gather:
            gather.gather(film)
        elif peer:
            dist.all_reduce(token)
        ev[k][2].record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches_per_frame = int(ctx.last_kernel_ms()[1])      # 1, or 1 + the long-ray round kernels of a partitioned frame
    total_ms = ev[0][0].elapsed_time(ev[-1][2])
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b, _ in ev]))
    tt = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms_max = float(tt[0]), float(tt[1])
    rays = W * H
    value = rays * args.steps / (total_ms * 1e-3) / 1e6
    if peer and rank == 0:
        api.memcpy(ctx, film.data_ptr(), shared.ptr, H * W * 16, 2)
    hits = int((film[..., :3].sum(dim=2) > 0).sum().item()) if rank == 0 else 0

    # ---- e2e: the call a user makes, host film in pinned memory, copies inside the timed region
    opts_sync = ctx.ls_opts(part=part)
    host = api.PinnedArray((H, W, 4), np.float32)
    host.array[...] = bg
    film_bytes = H * W * 16

    # multi-GPU e2e: ONE host film in POSIX shared memory, mapped and page-locked by every rank; each rank's kernels read the old
    # pixel of their misses and store the pixels they own over their own PCIe link -- the frame is assembled in host memory with
    # no gather and no staging copy.  (--gather nccl keeps the copy-in / NCCL gather / copy-out path.)
    host_shared = None
    if world > 1 and peer:
        def exchange_name(n):
            box = [n]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        host_shared = api.SharedHostFilm(H, W, rank, exchange_name)
        if rank == 0:
            host_shared.array[...] = bg
        dist.barrier()

    def e2e_step():
        if world == 1:
            ctx.render_levelset(grid, cam, sh, host.array, opts=opts_sync)       # pinned host film, read and written in place
        elif peer:
            ctx.render_levelset(grid, cam, sh, host_shared.array, opts=opts_sync)    # synchronous: returns when this rank's pixels are in the host film
            dist.barrier()                                                           # the frame is complete when every rank is done
        else:
            film.copy_(torch.from_numpy(host.array), non_blocking=True)          # H2D of this step's film
            step_opts = ctx.ls_opts(part=part)
            step_opts.flags |= abi.ASYNC
            ctx.render_levelset(grid, cam, sh, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=step_opts)
            gather.gather(film)
            if rank == 0:
                torch.from_numpy(host.array).copy_(film)                         # D2H of the finished frame
            torch.cuda.synchronize()

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = rays * args.steps / float(te[0]) / 1e6
    e2e_match = None
    if host_shared is not None:
        if rank == 0:
            e2e_match = bool(np.array_equal(host_shared.array, film.cpu().numpy()))     # same frame as the peer-store path
        dist.barrier()
        host_shared.close()

    upload = None
    if rank == 0 and world == 1 and not args.workload.startswith("c4"):
        # the one-time host -> device upload of the serialised grid (GridHandle::deviceUpload), reported separately
        hbuf = api.PinnedArray((grid.info.bytes,), np.uint8)
        ctx.L.vdbrt_grid_download(ctx.handle, grid.handle, hbuf.ptr, grid.info.bytes)
        t0 = time.perf_counter()
        g2 = ctx.upload(hbuf.array)
        ctx.synchronize()
        upload = {"bytes": int(grid.info.bytes), "ms": (time.perf_counter() - t0) * 1e3, "what": "vdbrt_upload_grid from pinned host memory incl. validation, the node-bbox kernel and the halo-block kernel (2944 B per leaf next to the grid)"}
        g2.free()
        hbuf.free()
    if rank == 0:
        counters = ctx.count_levelset(grid, cam).as_dict()      # separate instrumented launch, not timed
        bpr = bytes_per_ray(counters)
        peak, peak_src = measured_peak()
        # bytes the dominant kernel moves per launch on THIS rank: its share of the frame's rays
        rays_per_launch = rays / world
        achieved = bpr * rays_per_launch / (kernel_ms_max * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_levelset_c2_dram_bytes.json")
        if os.path.exists(tpath) and world == 1 and args.workload == "c2":  # ncu capture of exactly this workload
            try:
                traffic = json.load(open(tpath))["dram_bytes_per_launch"]
            except Exception:
                traffic = None
        line = {
            "metric": "primary Mrays/s, level-set ray tracer", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload + (": union of %d level-set spheres" % int(R) if args.workload.startswith("c4") else ": level-set torus R=%g r=%g" % (R, r))
                                   + " voxel 1 hw 3 (%d active voxels, %.2f GB grid, GPU-built in %.2f s), %dx%d, 1 spp, DiffuseShader, perspective camera"
                                   % (grid.info.active_voxels, grid.info.bytes / 1e9, build_s, W, H),
                       "partition": ("%d GPU(s), %dx%d tiles interleaved, " % (world, TILE_W, TILE_H)
                                     + ("render kernels store straight into rank 0's film over NVLink (CUDA IPC)" if peer else "NCCL gather to rank 0"))
                       if world > 1 else "single GPU",
                       "l2": "grid (%.2f GB) is larger than the 126 MB L2; no explicit flush" % (grid.info.bytes / 1e9),
                       "hit_pixels": hits},
            # single GPU: the pinned host film is read (old pixel of every miss) and written (every pixel) in place by the kernels;
            # multi GPU: the film is copied into and out of rank 0's shared device film
            "e2e": {"value": e2e_value, "unit": "Mrays/s",
                    "h2d_bytes_per_step": (W * H - hits) * 16 if (world == 1 or peer) else film_bytes, "d2h_bytes_per_step": film_bytes,
                    "ms_per_step": float(te[0]) * 1e3 / args.steps,
                    "how": ("vdbrt_render_levelset on a pinned host film (tools::Film): misses read their old pixel and all pixels are "
                            "stored over PCIe by the render kernel itself, no staging copies") if world == 1 else
                           ("one host film in shared memory, page-locked by every rank: each rank's kernels read / store its own pixels over "
                            "its own PCIe link, a barrier ends the frame; checked against rank 0's device-gathered frame") if peer else
                           "H2D of the film into rank 0's device film, partitioned render, NCCL gather, D2H of the frame",
                    "frame_matches_device_path": e2e_match},
            "gpu_launches": args.steps * launches_per_frame,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "kernel": "k_render_levelset" + (" + %d long-ray round kernels (k_long_scout/march/finish)" % (launches_per_frame - 1) if launches_per_frame > 1 else ""),
                         "kernel_ms": kernel_ms_max,
                         "algorithmic_bytes_per_ray": bpr, "rays_per_launch": rays_per_launch, "counters": counters},
            "clocks": clocks,
            "grid_upload": upload,
        }
        if not args.no_extras and world == 1 and args.workload == "c2":
            try:
                line["extras"] = extras(ctx, api, abi, torch)
            except Exception as e:
                line["extras"] = {"error": str(e)}
        if not args.no_cpu_baseline and world == 1 and not args.workload.startswith("c4"):
            try:
                line["cpu_baseline"] = cpu_reference(R, r, W, H, 3, 1, gpu_film=host.array.copy())
            except Exception as e:  # the checker is optional for the product arm
                line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": str(e)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
