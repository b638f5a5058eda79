# renders the c2 frame N times; `device` as the second argument keeps the film in device memory (what bench.py's `value` times),
# otherwise the film is a pinned host array (what `e2e` times)
import sys, numpy as np
sys.path.insert(0,'/root/repo')
from openvdb_b200 import api, _abi as abi
ctx=api.Context(0)
big=ctx.build_torus(650.0,325.0)
W,H=1920,1080
cam=api.vdb_render_camera(W,H,(0,1.5*650,3*(650+325.0)),(0,0,0))
sh=api.make_shader(abi.SHADER_DIFFUSE)
device = len(sys.argv) > 2 and sys.argv[2] == 'device'
if device:
    import torch
    dfilm = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
else:
    film=api.PinnedArray((H,W,4),np.float32); film.array[...]=(0,0,0,1)
for it in range(int(sys.argv[1]) if len(sys.argv)>1 else 3):
    if device:
        ctx.render_levelset(big,cam,sh,dfilm.data_ptr(),width=W,height=H,memspace=abi.MEM_DEVICE,opts=ctx.ls_opts(uniform_bg=True))
    else:
        ctx.render_levelset(big,cam,sh,film.array)
    print(ctx.last_kernel_ms())
