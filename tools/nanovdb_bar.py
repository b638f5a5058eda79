"""The prior-art bar: NanoVDB's own zeroCrossing kernels (oracle/nanovdb_bar.cu -> oracle/_ref/libnvbar.so,
compiled from the reference's headers) on the bench's grids and cameras, on the same GPU, next to this library's kernel.
A different algorithm (float rays, raw-voxel sign change): a stated bar, not a parity target.  Bench / test use only."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "oracle", "_ref", "libnvbar.so")


def run(workloads=("c2", "c4"), iters=10):
    import torch
    from openvdb_b200 import api, _abi as abi
    import bench
    if not os.path.exists(LIB):
        return {"unavailable": LIB + " not built (make -C oracle ref; needs /root/reference)"}
    L = C.CDLL(LIB)
    L.nvbar_render.argtypes = [C.c_void_p, C.POINTER(abi.Camera), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
    ctx = api.Context(torch.cuda.current_device())
    out = {"what": "nanovdb::math::zeroCrossing (float HDDA, the reference's own GPU path) with the launch shapes of ex_raytrace_level_set "
                   "(one thread per pixel, 512-thread blocks) and of renderIsoSurfacePersistentKernel (256 threads x 4 per SM, 32 pixels per warp ticket); "
                   "same grid, same vdb_render camera, 4-byte output per pixel; a different algorithm from the OpenVDB CPU tracer, no parity expected"}
    for name in workloads:
        wl = bench.WORKLOADS[name]
        W, H = wl["W"], wl["H"]
        g = bench.build_gpu_grid(ctx, api, name)
        cam = api.vdb_render_camera(W, H, *wl["cam"])
        film = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        ms = []
        for _ in range(4):
            ctx.render_levelset(g, cam, api.make_shader(abi.SHADER_DIFFUSE), film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE,
                                opts=ctx.ls_opts(uniform_bg=True))
            ms.append(ctx.last_kernel_ms()[0])
        mine_hits = int((film[..., :3].sum(dim=2) > 0).sum().item())
        del film
        host = g.download()
        g.free()
        dev = torch.from_numpy(host).cuda()           # the serialised NanoGrid<float>, as GridHandle::deviceUpload would place it
        assert dev.data_ptr() % 32 == 0
        res = {"vdbrt_ms": float(np.median(ms[1:])), "vdbrt_hit_pixels": mine_hits}
        for mode, key in ((0, "thread_per_pixel"), (1, "persistent")):
            t, h = C.c_float(), C.c_uint64()
            rc = L.nvbar_render(dev.data_ptr(), C.byref(cam), mode, 3, iters, C.byref(t), C.byref(h))
            res["nanovdb_%s_ms" % key] = t.value if rc == 0 else None
            res["nanovdb_%s_hit_pixels" % key] = int(h.value)
        res["hit_count_difference"] = res["nanovdb_persistent_hit_pixels"] - mine_hits
        best = min(v for k, v in res.items() if k.endswith("_ms") and k.startswith("nanovdb") and v)
        res["vdbrt_over_nanovdb_time"] = res["vdbrt_ms"] / best
        out[name] = res
        del dev
        torch.cuda.empty_cache()
    ctx.close()
    return out


if __name__ == "__main__":
    import json
    print(json.dumps(run(tuple(sys.argv[1:]) or ("c2",)), indent=1))
