#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric: primary Mrays/s on level-set & fog VDBs at 1/2/4/8 B200, ms per 4K frame.

Headline workload at every N ("c4", BASELINE config 4): union of 10 000 random level-set spheres (~1.03 G active voxels, 12.7 GB
NanoVDB grid built on the GPU by the library's own builder, replicated per GPU), 3840x2160, 1 spp, DiffuseShader, vdb_render's
perspective camera at (0, 0, 3*2048).  `ms_per_step` is the ms per 4K frame.  The same JSON line carries two secondary objects,
each with its own value / e2e / roofline / cpu_baseline:
  "c2"      level-set torus R=650 r=325 (50 M active voxels), 1920x1080, 1 spp, Diffuse          (BASELINE config 2)
  "c3_fog"  fog volume of a level-set sphere r=509 (1024^3 bbox), VolumeRender step 0.5, 1920x1080 (BASELINE config 3)

  value         whole-frame primary rays / device time, grid and film resident in HBM (CUDA events, max over ranks)
  e2e           the same metric through the C ABI with a pinned HOST film (tools::Film): the kernels read the old pixel of every
                miss and store every pixel over PCIe inside the timed region
  roofline      algorithmic bytes per ray (counted by an instrumented launch, SURVEY 8d formula) x rays / kernel time against
                the measured HBM peak; `traffic`, `l2_frac`, `issue_slot_util` are constants from the committed ncu capture of
                the same workload (profiles/), labelled as such
  cpu_baseline  the reference's own CPU path (oracle/_ref: unmodified OpenVDB, all host threads through a std::thread
                stand-in for TBB) on a bounded sample of the same workload, with the pixels that differ from the GPU's frame

N > 1 (torchrun, one rank per GPU): the grid is replicated, the film is split into 64x60-pixel tiles interleaved over the
ranks (vdbrt_partition), every rank's render kernels store the pixels they own straight into rank 0's film over NVLink
(CUDA IPC mapping) and a 4-byte all-reduce orders the streams ("strong" scaling: the frame is fixed).
`--impl reference` times the reference CPU implementation alone on the same workload (rank 0 only).
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

TILE_W, TILE_H = 64, 60
SEED = 20240607
WARMUP_FLOOR = 3

# kind, grid recipe, film, camera (translation, look-at), and the reference arm's bounded sample (film divided by `div` per axis)
WORKLOADS = {
    "c4": dict(kind="levelset", grid=("spheres", 10000, 1988.0), W=3840, H=2160, cam=((0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0)), div=4,
               text="c4: union of 10000 level-set spheres (mt19937_64 seed 20240607, centres U(-1988,1988)^3, radii U(10,60), voxel 1, half-width 3; "
                    "~1.03 G active voxels, 12.7 GB NanoVDB grid), 3840x2160, 1 spp, DiffuseShader, vdb_render perspective camera at (0,0,6144)"),
    "c2": dict(kind="levelset", grid=("torus", 650.0, 325.0), W=1920, H=1080, cam=((0.0, 1.5 * 650, 3.0 * 975), (0.0, 0.0, 0.0)), div=1,
               text="c2: level-set torus R=650 r=325 voxel 1 half-width 3 (50.0 M active voxels, 0.58 GB NanoVDB grid), 1920x1080, 1 spp, "
                    "DiffuseShader, vdb_render perspective camera at (0,975,2925)"),
    "c3": dict(kind="fog", grid=("fogsphere", 509.0), W=1920, H=1080, cam=((0.0, 0.0, 3 * 509.0), (0.0, 0.0, 0.0)), div=2,
               text="c3: fog volume (sdfToFogVolume of a level-set sphere r=509, 1024^3 bbox, 0.19 GB NanoVDB grid), VolumeRender absorption 0.1 "
                    "scattering 1.5, primary step 0.5, shadow step 3, 1920x1080, vdb_render perspective camera at (0,0,1527)"),
    # quick functional checks (CPU contract tests, smoke runs) -- not bench lines
    "c4-small": dict(kind="levelset", grid=("spheres", 300, 600.0), W=1280, H=720, cam=((0.0, 0.0, 3 * 660.0), (0.0, 0.0, 0.0)), div=2,
                     text="c4-small: union of 300 level-set spheres (extent 600), 1280x720, 1 spp, DiffuseShader"),
    "c4-tiny": dict(kind="levelset", grid=("spheres", 40, 200.0), W=320, H=180, cam=((0.0, 0.0, 3 * 260.0), (0.0, 0.0, 0.0)), div=1,
                    text="c4-tiny: union of 40 level-set spheres (extent 200), 320x180, 1 spp, DiffuseShader"),
    "c2-small": dict(kind="levelset", grid=("torus", 160.0, 80.0), W=640, H=360, cam=((0.0, 240.0, 720.0), (0.0, 0.0, 0.0)), div=1,
                     text="c2-small: level-set torus R=160 r=80, 640x360, 1 spp, DiffuseShader"),
    "c3-small": dict(kind="fog", grid=("fogsphere", 60.0), W=320, H=180, cam=((0.0, 0.0, 180.0), (0.0, 0.0, 0.0)), div=1,
                     text="c3-small: fog volume of a level-set sphere r=60, step 0.5, 320x180"),
}
SECONDARY = {"c4": ("c2", "c3"), "c4-small": ("c2-small", "c3-small"), "c4-tiny": ("c2-small", "c3-small")}
METRIC = "primary Mrays/s, level-set ray tracer (ms_per_step = ms per frame of the workload; c4: ms per 4K frame)"


def config_for(name, gpus):
    """identical in the GPU arm and the reference arm: the driver compares the two"""
    return {"workload": WORKLOADS[name]["text"],
            "partition": "single GPU" if gpus <= 1 else "%d GPUs, grid replicated, %dx%d film tiles interleaved over the ranks" % (gpus, TILE_W, TILE_H),
            "l2": "grid larger than the 126 MB L2, no explicit flush" if not name.startswith("c3") and "small" not in name and "tiny" not in name
                  else "grid not larger than the L2 on this small / fog workload; no explicit flush (every frame re-reads it through the same caches)",
            "warmup_floor": WARMUP_FLOOR,
            "tile_order": "GPU arm: level-set frames after the first of a sequence hand out the heaviest 8x4 tiles first, from the previous frame's measured "
                          "tile costs (a hint for the ORDER of the work queue only: every ray of every frame is traced; `without_tile_cost_history` is the same "
                          "frame in plain tile order)"}


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


def bytes_per_ray(c, kind):
    """B_LS = 32 n_R + 16 n_U + 16 n_L + 12 n_V + 32 n_S + 16;  B_fog = 32 n_R + 16 (n_U + n_L) + 96 (n_P + n_Sh) + 16"""
    n = float(c["rays"])
    if kind == "fog":
        return (32 * c["root_probes"] + 16 * (c["upper_probes"] + c["lower_probes"]) + 96 * (c["primary_samples"] + c["shadow_samples"])) / n + 16.0
    return (32 * c["root_probes"] + 16 * c["upper_probes"] + 16 * c["lower_probes"] + 12 * c["voxel_probes"] + 32 * c["stencil_refills"]) / n + 16.0


def profile_constants(name):
    """what only a profiler sees (DRAM bytes per launch, L2 throughput, issue slots): constants of the committed ncu capture of this
    workload's kernel (profiles/r02_<workload>_ncu_constants.json), NOT measured in this run"""
    p = os.path.join(ROOT, "profiles", "r02_%s_ncu_constants.json" % name)
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return None


# ---------------------------------------------------------------------------------------------------------------------
# the reference's own CPU implementation (oracle/_ref) -- checker and baseline, never the product
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference(name, steps, warmup, gpu_frame=None):
    """unmodified OpenVDB LevelSetRayTracer / VolumeRender, threaded, on a bounded sample of the workload: the same grid and camera
    at (W/div) x (H/div) pixels.  gpu_frame(w, h) -> the GPU's frame of the same sample (for the parity report).
    Falls back to the oracle port when oracle/_ref is not built."""
    from tests import refapi
    from openvdb_b200 import api, _abi as abi
    wl = WORKLOADS[name]
    w, h = max(wl["W"] // wl["div"], 1), max(wl["H"] // wl["div"], 1)
    tr, look = wl["cam"]
    cores = os.cpu_count() or 1
    fog = wl["kind"] == "fog"
    t0 = time.perf_counter()
    if os.path.exists(refapi.REF_SO):
        ref = refapi.Ref()
        ref.set_threads(cores)
        recipe = wl["grid"]
        if recipe[0] == "torus":
            g = ref.torus(recipe[1], recipe[2])
        elif recipe[0] == "spheres":
            g = ref.spheres_union_mt(api.random_spheres(int(recipe[1]), SEED, recipe[2], 10.0, 60.0), cores)
        else:
            ls = ref.sphere(recipe[1])
            g = ref.fog_from_levelset(ls)
            ref.free(ls)
        build_s = time.perf_counter() - t0
        d = refapi.camera_desc(w, h, translation=tr, lookat=look)
        film = refapi.new_film(w, h)
        times = []
        if fog:
            vo = ref.vol_defaults()
            vo.primary_step = 0.5
            for it in range(warmup + steps):
                t = ref.render_volume(g, d, vo, film, threaded=True)
                if it >= warmup:
                    times.append(t)
        else:
            sh = refapi.shader(abi.SHADER_DIFFUSE)
            for it in range(warmup + steps):
                film[...] = (0, 0, 0, 1)
                t = ref.render_levelset(g, d, sh, film, threaded=True)
                if it >= warmup:
                    times.append(t)
        ref.free(g)
        kind = "reference"
    else:
        oracle = refapi.Oracle()
        ctx = api.Context(0)
        grid = build_gpu_grid(ctx, api, name)
        og = oracle.open(grid.download())
        build_s = time.perf_counter() - t0
        cam = api.vdb_render_camera(w, h, tr, look)
        film = refapi.new_film(w, h)
        times = []
        for it in range(warmup + steps):
            t1 = time.perf_counter()
            if fog:
                vo = api.vol_opts_default()
                vo.primary_step = 0.5
                oracle.render_volume(og, cam, vo, film, threads=cores)
            else:
                film[...] = (0, 0, 0, 1)
                oracle.render_levelset(og, cam, api.make_shader(abi.SHADER_DIFFUSE), film, threads=cores)
            if it >= warmup:
                times.append(time.perf_counter() - t1)
        kind = "port"
    t = float(np.median(times))
    hits = int((film[..., 3] > 0).sum()) if fog else int((film[..., :3].sum(axis=2) > 0).sum())
    parity = None
    if gpu_frame is not None:
        try:
            gf = gpu_frame(w, h)
            diff = np.abs(gf.astype(np.float64) - film.astype(np.float64))
            if fog:
                # fog: exp() is CUDA's on the GPU and glibc's on the host -> tolerance 1e-4 rel / 1e-3 abs (north_star), alpha>0 mask exact
                bad = int((diff > 1e-3 + 1e-4 * np.abs(film)).any(axis=2).sum())
                mask = int(((gf[..., 3] > 0) != (film[..., 3] > 0)).sum())
                parity = {"against": kind, "pixels": int(w * h), "mismatched_pixels": bad, "tolerance": "1e-4 rel + 1e-3 abs",
                          "alpha_mask_mismatches": mask, "bit_identical_pixels": int((gf == film).all(axis=2).sum()), "max_abs_diff": float(diff.max())}
            else:
                parity = {"against": kind, "pixels": int(w * h), "mismatched_pixels": int((gf != film).any(axis=2).sum()),
                          "tolerance": "bit-exact", "max_abs_diff": float(diff.max())}
        except Exception as e:
            parity = {"error": str(e)}
    return {"value": w * h / t / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind, "parity": parity,
            "threads": "%d std::thread workers behind the TBB stand-in of oracle/tbb_shim (not oneTBB: this image ships none)" % cores,
            "sample": "%dx%d pixels of the same grid and camera (1/%d of the frame's rays), median of %d runs after %d warm-up, %d %s pixels; "
                      "grid built by the reference in %.1f s (not timed)" % (w, h, wl["div"] ** 2, steps, warmup, hits, "alpha>0" if fog else "hit", build_s),
            "ms_per_sample": t * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference(args.workload, max(args.steps, 1), max(min(args.warmup, 2), 1))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_sample"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_for(args.workload, args.gpus), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def build_gpu_grid(ctx, api, name):
    recipe = WORKLOADS[name]["grid"]
    if recipe[0] == "torus":
        return ctx.build_torus(recipe[1], recipe[2])
    if recipe[0] == "spheres":
        return ctx.build_spheres(api.random_spheres(int(recipe[1]), SEED, recipe[2], 10.0, 60.0))
    ls = ctx.build_sphere(recipe[1])
    fog = ctx.build_fog(ls)
    ls.free()
    return fog


class Rig:
    """what every measurement of one process shares: context, stream, ranks"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from openvdb_b200 import api, _abi as abi
        self.torch, self.dist, self.api, self.abi = torch, dist, api, abi
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: libvdbrt.so has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.ctx = api.Context(self.local)
        # one non-default torch stream carries the library's kernels, torch's copies, NCCL and the timing events
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        assert self.stream.cuda_stream != 0
        self.ctx.set_stream(self.stream.cuda_stream)
        self.token = torch.zeros(1, dtype=torch.float32, device="cuda")
        self.args = args

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()


def measure(rig, name, steps, warmup, sampler=None):
    """value / e2e / roofline of one workload on all ranks of this job; returns (dict, grid-less state for the parity frame)"""
    torch, dist, api, abi, ctx = rig.torch, rig.dist, rig.api, rig.abi, rig.ctx
    world, rank = rig.world, rig.rank
    wl = WORKLOADS[name]
    W, H, fog = wl["W"], wl["H"], wl["kind"] == "fog"
    tr, look = wl["cam"]
    t0 = time.perf_counter()
    grid = build_gpu_grid(ctx, api, name)              # replicated on every GPU
    ctx.synchronize()
    build_s = time.perf_counter() - t0
    cam = api.vdb_render_camera(W, H, tr, look)
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    part = api.partition(rank, world, TILE_W, TILE_H) if world > 1 else None
    film = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    bg = (0.0, 0.0, 0.0, 1.0)                           # a fresh tools::Film (RayTracer.h:235)
    if fog:
        vo = api.vol_opts_default()
        vo.primary_step = 0.5
        if part is not None:
            vo.part = part
        vo.flags |= abi.ASYNC
    else:
        opts = ctx.ls_opts(part=part, uniform_bg=True)
        opts.flags |= abi.ASYNC

    # multi-GPU frame assembly: every rank's render kernel stores the tiles it owns straight into rank 0's film (CUDA IPC mapping,
    # NVLink peer stores) and a 4-byte all-reduce orders the streams -- compute and gather are one kernel.
    film_ptr = film.data_ptr()
    shared = None
    if world > 1:
        def exchange(h):
            t = torch.from_numpy(h.copy()).cuda()
            dist.broadcast(t, src=0)
            return t.cpu().numpy()
        shared = api.SharedFilm(ctx, H, W, rank, exchange)
        film_ptr = shared.ptr

    def render():
        if fog:
            ctx.render_volume(grid, cam, vo, film_ptr, width=W, height=H, memspace=abi.MEM_DEVICE)
        else:
            ctx.render_levelset(grid, cam, sh, film_ptr, width=W, height=H, memspace=abi.MEM_DEVICE, bg=bg, opts=opts)

    for _ in range(warmup):
        render()
        if world > 1:
            dist.all_reduce(rig.token)
    rig.barrier()
    if sampler is not None:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        ev[k][0].record(rig.stream)
        render()
        ev[k][1].record(rig.stream)
        if world > 1:
            dist.all_reduce(rig.token)      # stream-ordered: rank 0 continues only after every rank's kernel has stored its tiles
        ev[k][2].record(rig.stream)
    rig.barrier()
    clocks = sampler.stop() if sampler is not None else None
    launches_per_frame = int(ctx.last_kernel_ms()[1])
    total_ms = ev[0][0].elapsed_time(ev[-1][2])
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b, _ in ev]))
    tt = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms_max = float(tt[0]), float(tt[1])
    rays = W * H
    value = rays * steps / (total_ms * 1e-3) / 1e6
    # the same frame without the tile-cost history of the previous frame (plain tile order), for the record
    plain = None
    if not fog:
        ctx.set_tuning(ls_history=0)
        pm = []
        for _ in range(4):
            render()
            if world > 1:
                dist.all_reduce(rig.token)
            rig.barrier()
            pm.append(ctx.last_kernel_ms()[0])
        ctx.set_tuning(ls_history=1)
        tp = torch.tensor([float(np.median(pm[1:]))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        plain = {"kernel_ms": float(tp[0]), "value": rays / (float(tp[0]) * 1e-3) / 1e6}
    if shared is not None and rank == 0:
        api.memcpy(ctx, film.data_ptr(), shared.ptr, H * W * 16, 2)
        ctx.synchronize()
    hits = 0
    if rank == 0:
        hits = int((film[..., 3] > 0).sum().item()) if fog else int((film[..., :3].sum(dim=2) > 0).sum().item())

    # ---- e2e: the call a user makes, host film in pinned memory, PCIe traffic inside the timed region
    host = api.PinnedArray((H, W, 4), np.float32) if world == 1 else None
    host_shared = None
    if world > 1:
        # ONE host film in POSIX shared memory, mapped and page-locked by every rank; each rank's kernels read the old pixel of
        # their misses and store the pixels they own over their own PCIe link -- assembled in host memory, no gather, no staging
        def exchange_name(n):
            box = [n]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        host_shared = api.SharedHostFilm(H, W, rank, exchange_name)
        if rank == 0:
            host_shared.array[...] = bg
        dist.barrier()
    else:
        host.array[...] = bg
    harr = host.array if world == 1 else host_shared.array
    if fog:
        vo_sync = api.vol_opts_default()
        vo_sync.primary_step = 0.5
        if part is not None:
            vo_sync.part = part
    else:
        opts_sync = ctx.ls_opts(part=part)

    def e2e_step():
        if fog:
            ctx.render_volume(grid, cam, vo_sync, harr)          # synchronous: returns when this rank's pixels are in the host film
        else:
            ctx.render_levelset(grid, cam, sh, harr, opts=opts_sync)
        if world > 1:
            dist.barrier()                                        # the frame is complete when every rank is done

    e2e_steps = steps
    for _ in range(WARMUP_FLOOR):
        e2e_step()
    rig.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    rig.barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = rays * e2e_steps / float(te[0]) / 1e6
    e2e_match = None
    if rank == 0:
        e2e_match = bool(np.array_equal(harr, film.cpu().numpy()))     # the host frame is the device frame, bit for bit
    if host_shared is not None:
        dist.barrier()
        host_shared.close()
    if host is not None:
        host.free()
    if shared is not None:
        rig.barrier()
        shared.close()

    out = None
    if rank == 0:
        counters = (ctx.count_volume(grid, cam, vo_sync_plain(api)) if fog else ctx.count_levelset(grid, cam)).as_dict()   # instrumented launch, not timed
        bpr = bytes_per_ray(counters, wl["kind"])
        peak, peak_src = measured_peak()
        rays_per_launch = rays / world                  # this rank's share of the frame's rays
        achieved = bpr * rays_per_launch / (kernel_ms_max * 1e-3) / 1e9
        pc = profile_constants(name) if world == 1 else None
        kernel = "k_fog_shadow" if fog else "k_render_levelset"
        film_bytes = H * W * 16
        out = {
            "workload": wl["text"], "metric": "primary Mrays/s", "value": value, "unit": "Mrays/s", "ms_per_step": total_ms / steps,
            "kernel_ms": kernel_ms_max, "sync_ms": total_ms / steps - kernel_ms_max, "gpu_launches_per_step": launches_per_frame,
            ("alpha_pixels" if fog else "hit_pixels"): hits, "without_tile_cost_history": plain,
            "grid": {"active_voxels": int(grid.info.active_voxels), "bytes": int(grid.info.bytes), "leaves": int(grid.info.leaf_count), "gpu_build_s": build_s},
            "e2e": {"value": e2e_value, "unit": "Mrays/s",
                    "h2d_bytes_per_step": (0 if fog else (W * H - hits) * 16) + C.sizeof(abi.Camera) + (C.sizeof(abi.VolOpts) if fog else C.sizeof(abi.LsOpts) + C.sizeof(abi.Shader)),
                    "d2h_bytes_per_step": film_bytes, "ms_per_step": float(te[0]) * 1e3 / e2e_steps,
                    "how": ("vdbrt_render_%s on a pinned host film (tools::Film): %sall pixels are stored over PCIe by the render kernel itself, "
                            "no staging copies; h2d = old pixels of the misses + the camera / options PODs of the call"
                            % ("volume" if fog else "levelset", "" if fog else "misses read their old pixel and ")) if world == 1 else
                           "one host film in shared memory, page-locked by every rank: each rank's kernels read / store its own pixels over its own PCIe link, a barrier ends the frame",
                    "frame_matches_device_path": e2e_match},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": pc.get("dram_bytes_per_launch") if pc else None,
                         "traffic_source": (pc.get("source") if pc else None),
                         "l2_frac": pc.get("l2_throughput_frac") if pc else None, "issue_slot_util": pc.get("issue_slot_util") if pc else None,
                         "active_lanes": pc.get("active_lanes") if pc else None,
                         "peak_source": peak_src,
                         "kernel": kernel + ((" (+ %d more launches per frame: k_fog_primary / k_fog_resolve of the wavefront)" if fog else " (+ %d more launches per frame: probe / long-ray round kernels)") % (launches_per_frame - 1) if launches_per_frame > 1 else ""),
                         "kernel_ms": kernel_ms_max, "algorithmic_bytes_per_ray": bpr, "rays_per_launch": rays_per_launch, "counters": counters},
        }
        if clocks is not None:
            out["clocks"] = clocks
    state = dict(grid=grid, cam_args=(tr, look), fog=fog)
    return out, state


def vo_sync_plain(api):
    vo = api.vol_opts_default()
    vo.primary_step = 0.5
    return vo


def gpu_sample_frame(rig, state, w, h):
    """the GPU's frame of the reference arm's bounded sample (same grid, same camera, w x h pixels), on the host"""
    api, abi, ctx = rig.api, rig.abi, rig.ctx
    from tests import refapi
    cam = api.vdb_render_camera(w, h, *state["cam_args"])
    film = refapi.new_film(w, h)
    if state["fog"]:
        ctx.render_volume(state["grid"], cam, vo_sync_plain(api), film)
    else:
        ctx.render_levelset(state["grid"], cam, api.make_shader(abi.SHADER_DIFFUSE), film)
    return film


def c5_share(rig):
    """informational: C5 (the C4 union under its own fog volume, 16 jittered samples per pixel, fog.over(level set)) names the whole
    8-GPU box; this times rank 0's share of an 8-way tile split on one GPU, once"""
    torch, api, abi, ctx = rig.torch, rig.api, rig.abi, rig.ctx
    W, H, spp, share = 3840, 2160, 16, 8
    g = build_gpu_grid(ctx, api, "c4")
    try:
        fog = ctx.build_fog(g)
        cam = api.vdb_render_camera(W, H, *WORKLOADS["c4"]["cam"])
        part = api.partition(0, share, TILE_W, TILE_H)
        film = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        f2 = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        sampler = ClockSampler(rig.local)
        sampler.start()
        calls = []
        for it in range(2):               # the first call also allocates the wavefront's record buffer; the second one is reported
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
            calls.append((ms_ls, ms_fog, ms_over))
        clocks = sampler.stop()
        ms_ls, ms_fog, ms_over = calls[-1]
        rays = W * H * spp // share
        out = {"ms_level_set": ms_ls, "ms_fog": ms_fog, "ms_over_whole_film": ms_over, "ms_per_frame": ms_ls + ms_fog + ms_over,
               "first_call_ms": {"level_set": calls[0][0], "fog": calls[0][1]}, "clocks": clocks,
               "primary_rays": 2 * rays, "Mrays_per_s": 2 * rays / (ms_ls + ms_fog + ms_over) / 1e3, "fog_grid_gb": fog.info.bytes / 1e9}
        fog.free()
    except Exception as e:
        out = {"error": str(e)}
    g.free()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vdbrt", choices=["vdbrt", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the c2 / c3_fog objects")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational timings (C5 share, NanoVDB example kernels)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    rig = Rig(args)
    dist, world, rank = rig.dist, rig.world, rig.rank
    warmup = max(args.warmup, WARMUP_FLOOR)          # the timing rules ask for at least 3 warm-up steps; declared in config.warmup_floor
    sampler = ClockSampler(rig.local) if rank == 0 else None
    head, state = measure(rig, args.workload, args.steps, warmup, sampler)
    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config_for(args.workload, world), "e2e": head["e2e"],
                "gpu_launches": args.steps * head["gpu_launches_per_step"], "roofline": head["roofline"], "clocks": head.get("clocks"),
                "kernel_ms": head["kernel_ms"], "sync_ms": head["sync_ms"], "grid": head["grid"], "without_tile_cost_history": head["without_tile_cost_history"], "hit_pixels": head.get("hit_pixels", head.get("alpha_pixels"))}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_reference(args.workload, 3, 1, gpu_frame=lambda w, h: gpu_sample_frame(rig, state, w, h))
            except Exception as e:  # the checker is optional for the product arm
                line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": str(e)}
    state["grid"].free()
    if not args.no_secondary:
        for sec in SECONDARY.get(args.workload, ()):
            res, st = measure(rig, sec, max(args.steps // 2, 3), warmup)
            if rank == 0:
                if world == 1 and not args.no_cpu_baseline:
                    try:
                        res["cpu_baseline"] = cpu_reference(sec, 3, 1, gpu_frame=lambda w, h: gpu_sample_frame(rig, st, w, h))
                    except Exception as e:
                        res["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": str(e)}
                line["c3_fog" if WORKLOADS[sec]["kind"] == "fog" else sec.split("-")[0]] = res
            st["grid"].free()
    if rank == 0 and world == 1 and not args.no_extras and args.workload == "c4":
        ex = {}
        try:
            ex["c5_overlay_4k_16spp_rank0_of_8"] = c5_share(rig)
        except Exception as e:
            ex["c5_overlay_4k_16spp_rank0_of_8"] = {"error": str(e)}
        try:
            from tools import nanovdb_bar
            ex["nanovdb_example_kernels"] = nanovdb_bar.run()
        except Exception as e:
            ex["nanovdb_example_kernels"] = {"unavailable": str(e)}
        line["extras"] = ex
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
