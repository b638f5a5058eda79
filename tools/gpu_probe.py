# first GPU parity probe: prints mismatch statistics (CUDA vs oracle vs reference)
import sys, time, numpy as np, ctypes as C
sys.path.insert(0,'/root/repo')
from tests.refapi import *
from openvdb_b200 import api, _abi as abi
R=Ref(); O=Oracle(); ctx=api.Context(0)
def cmp_ls(name, g, W, H, tr, look, shk=abi.SHADER_DIFFUSE, spp=1):
    buf=R.nanovdb(g); og=O.open(buf); dg=ctx.upload(buf)
    i=dg.info; oi=O.info(og)
    print(name,'node bbox gpu',list(i.node_bbox),'oracle',list(oi.node_bbox), 'leaves',i.leaf_count)
    d=camera_desc(W,H,translation=tr,lookat=look); cam=R.camera_pod(d)
    sh=api.make_shader(shk)
    f_o=new_film(W,H); t0=time.time(); aux_o,_=O.render_levelset(og,cam,sh,f_o,aux=True,spp=spp,jitter=api.jitter_table(0),threads=8); to=time.time()-t0
    f_g=new_film(W,H); aux=AuxArrays(W,H); pod=aux.pod()
    ctx.render_levelset(dg,cam,sh,f_g,aux=pod,spp=spp)
    ms,_=ctx.last_kernel_ms()
    print(name,'oracle s %.3f gpu kernel ms %.3f'%(to,ms),'hits',aux.hit.sum(), aux_o.hit.sum())
    print('  hit mismatches',(aux.hit!=aux_o.hit).sum(),'ijk',(aux.ijk!=aux_o.ijk).any(axis=1).sum(),'t_index',(aux.t_index!=aux_o.t_index).sum(),
          't_world',(aux.t_world!=aux_o.t_world).sum(),'xyz',(aux.xyz!=aux_o.xyz).any(axis=1).sum(),'nml',(aux.nml!=aux_o.nml).any(axis=1).sum(),
          'film',(f_g!=f_o).any(axis=2).sum(),'maxdiff',np.abs(f_g-f_o).max())
    f_g2=new_film(W,H); ctx.render_levelset(dg,cam,sh,f_g2,spp=spp); print('  no-aux film equal',np.array_equal(f_g,f_g2), 'ms',ctx.last_kernel_ms()[0])
    c=ctx.count_levelset(dg,cam,spp=spp); print('  counters',c.as_dict())
    return dg,og,cam
g=R.sphere(100); dg,og,cam=cmp_ls('C1-512',g,512,512,(0,0,300),(0,0,0))
cmp_ls('C1-1024',g,1024,1024,(0,0,300),(0,0,0))
cmp_ls('C1-spp4',g,256,256,(0,0,300),(0,0,0),spp=4)
cmp_ls('C1-normal',g,333,217,(120,80,260),(0,0,0),shk=abi.SHADER_NORMAL)
t=R.torus(60,25); cmp_ls('torus',t,640,360,(0,90,255),(0,0,0))
# fog
fg=R.fog_from_levelset(g); buf=R.nanovdb(fg); ofg=O.open(buf); dfg=ctx.upload(buf)
W=H=256; d=camera_desc(W,H,translation=(0,0,300),lookat=(0,0,0)); cam=R.camera_pod(d)
vo=api.vol_opts_default(); vo.primary_step=0.5
f_o=new_film(W,H); t0=time.time(); O.render_volume(ofg,cam,vo,f_o,threads=8); to=time.time()-t0
f_g=new_film(W,H); ctx.render_volume(dfg,cam,vo,f_g); ms,_=ctx.last_kernel_ms()
print('fog oracle s %.3f gpu ms %.3f'%(to,ms),'alpha>0 mism',((f_g[...,3]>0)!=(f_o[...,3]>0)).sum(),'exact-equal px',(f_g==f_o).all(axis=2).mean(),'maxabs',np.abs(f_g-f_o).max(),
      'maxrel',(np.abs(f_g-f_o)/np.maximum(np.abs(f_o),1e-6)).max())
print('  counters',ctx.count_volume(dfg,cam,vo).as_dict())
rays=R.camera_rays(d,[(i,j) for j in range(0,H,8) for i in range(0,W,8)])
s1,c1=O.volume_spans(ofg,rays); s2,c2=ctx.volume_spans(dfg,rays); print('spans equal',np.array_equal(s1,s2),np.array_equal(c1,c2))
h1=O.intersect(og,rays); h2=hits_to_dict(ctx.intersect(dg,rays),len(rays)); print('intersect equal', h1.tobytes()==h2.tobytes())
