# C2 torus as float, Fp8 and Fp16 grids (the reference's createNanoGrid<FloatGrid, FpX>), each rendered from its codes (native) and
# expanded: resident bytes and frame time.  Under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:k_render_levelset`
# the launches appear in the order printed here.
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from openvdb_b200 import api, _abi as abi
from tests import refapi
ctx = api.Context(0)
ref = refapi.Ref()
R, r = (650.0, 325.0) if len(sys.argv) < 2 else (float(sys.argv[1]), float(sys.argv[1]) / 2)
W, H = 1920, 1080
cam = api.vdb_render_camera(W, H, (0, 1.5 * R, 3 * (R + r)), (0, 0, 0))
sh = api.make_shader(abi.SHADER_DIFFUSE)
film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
t0 = time.time(); ls = ref.torus(R, r); print("reference torus built in %.1f s" % (time.time() - t0), flush=True)
frames = {}
for name, gtype in (("float", 1), ("fp8", 14), ("fp16", 15)):
    buf = ref.nanovdb(ls) if gtype == 1 else ref.nanovdb_quantized(ls, gtype)
    for native in ((1,) if gtype == 1 else (1, 0)):
        ctx.set_tuning(quant_native=native, ls_history=1)
        g = ctx.upload(buf)
        ms = []
        for it in range(6):
            ctx.render_levelset(g, cam, sh, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=ctx.ls_opts(uniform_bg=True, rounds=False))
            ms.append(ctx.last_kernel_ms()[0])
        f = film.cpu().numpy().copy()
        key = name
        same = np.array_equal(frames.setdefault(key, f), f)
        print("%-6s %-8s source %7.1f MB  resident %7.1f MB  frame %.3f ms (min of 5 after the first)  hits %d  same frame as the other mode: %s"
              % (name, "native" if (native and gtype != 1) else ("" if gtype == 1 else "expanded"), buf.size / 1e6, g.info.resident_bytes / 1e6, min(ms[1:]),
                 int((f[..., :3].sum(axis=2) > 0).sum()), same), flush=True)
        g.free()
