# times the fog wavefront on C3 and on a 1/64 share of C5's fog pass with every experiment build under tools/_variants/fog_*.so
import glob, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
names = sys.argv[1:] or sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(root, "tools/_variants/fog_*.so")))
code = r'''
import sys, numpy as np, torch
sys.path.insert(0, "%s")
from openvdb_b200 import api, _abi as abi
ctx = api.Context(0)
def go(name, fog, cam, W, H, spp, part, reps):
    film = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    vo = api.vol_opts_default(spp=spp, seed=0); vo.primary_step = 0.5
    if part is not None: vo.part = part
    ms = []
    for it in range(reps):
        ctx.render_volume(fog, cam, vo, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE); ms.append(ctx.last_kernel_ms()[0])
    print("%%-8s min %%.3f ms  alpha sum %%.3f" %% (name, min(ms), float(film[..., 3].sum().item())), flush=True)
ls = ctx.build_sphere(509.0); fog = ctx.build_fog(ls); ls.free()
go("c3", fog, api.vdb_render_camera(1920, 1080, (0, 0, 3 * 509.0), (0, 0, 0)), 1920, 1080, 1, None, 4)
fog.free()
g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0)); fog = ctx.build_fog(g); g.free()
go("c5 1/64", fog, api.vdb_render_camera(3840, 2160, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0)), 3840, 2160, 16, api.partition(0, 64, 64, 60), 2)
''' % root
for n in names:
    env = dict(os.environ, VDBRT_LIBRARY=os.path.join(root, "tools/_variants", n + ".so"))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, stdin=subprocess.DEVNULL)
    print(n, "|", " | ".join(r.stdout.strip().splitlines()) or r.stderr[-400:], flush=True)
