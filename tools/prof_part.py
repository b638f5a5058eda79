import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from openvdb_b200 import api, _abi as abi
ctx=api.Context(0)
big=ctx.build_torus(650.0,325.0)
W,H=1920,1080
cam=api.vdb_render_camera(W,H,(0,1.5*650,3*(650+325.0)),(0,0,0))
sh=api.make_shader(abi.SHADER_DIFFUSE)
film=torch.zeros((H,W,4),dtype=torch.float32,device='cuda')
def run(part,label):
    o=ctx.ls_opts(part=part,uniform_bg=True)
    ts=[]
    for it in range(4):
        ctx.render_levelset(big,cam,sh,film.data_ptr(),width=W,height=H,memspace=abi.MEM_DEVICE,opts=o)
        ts.append(ctx.last_kernel_ms()[0])
    print(label,'ms',['%.3f'%t for t in ts])
run(None,'full')
for n in (2,4,8):
    for tw,th in ((64,60),(64,64),(32,30),(128,120),(16,12)):
        t=[]
        for r in range(n):
            o=ctx.ls_opts(part=api.partition(r,n,tw,th),uniform_bg=True)
            ctx.render_levelset(big,cam,sh,film.data_ptr(),width=W,height=H,memspace=abi.MEM_DEVICE,opts=o)
            ctx.render_levelset(big,cam,sh,film.data_ptr(),width=W,height=H,memspace=abi.MEM_DEVICE,opts=o)
            t.append(ctx.last_kernel_ms()[0])
        print('n',n,'tile',tw,th,'per-rank ms',['%.3f'%x for x in t],'max %.3f'%max(t))
