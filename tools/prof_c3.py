import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from openvdb_b200 import api, _abi as abi
ctx=api.Context(0)
r=float(sys.argv[2]) if len(sys.argv)>2 else 509.0
t0=time.time(); ls=ctx.build_sphere(r); fog=ctx.build_fog(ls); ctx.synchronize(); print('build s',time.time()-t0,'fog bytes',fog.info.bytes,'leaves',fog.info.leaf_count)
W,H=(1920,1080) if r>200 else (512,512)
cam=api.vdb_render_camera(W,H,(0,0,3*r),(0,0,0))
vo=api.vol_opts_default(); vo.primary_step=0.5
film=api.PinnedArray((H,W,4),np.float32)
for it in range(int(sys.argv[1]) if len(sys.argv)>1 else 3):
    t0=time.time(); ctx.render_volume(fog,cam,vo,film.array); te=time.time()-t0
    print('kernel ms',ctx.last_kernel_ms(),'e2e ms',te*1e3,'alpha sum',float(film.array[...,3].sum()))
c=ctx.count_volume(fog,cam,vo).as_dict(); print(c)
n=c['rays']; print('B/ray', (32*c['root_probes']+16*(c['upper_probes']+c['lower_probes'])+96*(c['primary_samples']+c['shadow_samples']))/n+16)
