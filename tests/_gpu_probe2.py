# builders vs reference + C2-scale timing
import sys, time, numpy as np, ctypes as C
sys.path.insert(0,'/root/repo')
from tests.refapi import *
from openvdb_b200 import api, _abi as abi
R=Ref(); O=Oracle(); ctx=api.Context(0)
def compare(name, mine, refh, nprobe=400000):
    buf=mine.download(); i=mine.info
    st=R.stats(refh)
    print(name,'mine: leaves',i.leaf_count,'lower',i.lower_count,'upper',i.upper_count,'voxels',i.active_voxels,'nodebbox',list(i.node_bbox),'idx',list(i.index_bbox))
    print(name,'ref : ',st)
    rb=R.nanovdb(refh); og=O.open(rb); oi=O.info(og); print(name,'ref nano: leaves',oi.leaf_count,'lower',oi.lower_count,'upper',oi.upper_count,'bytes',rb.size,'mine bytes',buf.size)
    mh=R.from_nanovdb(buf); st2=R.stats(mh); print(name,'mine->openvdb',st2)
    rng=np.random.default_rng(0)
    nb=np.array(st['node_bbox'])
    ijk=rng.integers(nb[:3]-20, nb[3:]+20, size=(nprobe,3)).astype(np.int32)
    v1,a1=R.probe(refh,ijk); v2,a2=R.probe(mh,ijk)
    print(name,'random probes: value mismatches',(v1!=v2).sum(),'active mismatches',(a1!=a2).sum(), 'active frac',a1.mean())
    # near-surface probes: take active voxels and their neighbours
    act=ijk[a1>0][:50000]
    nbr=(act[:,None,:]+rng.integers(-2,3,size=(len(act),4,3))).reshape(-1,3).astype(np.int32)
    v1,a1=R.probe(refh,nbr); v2,a2=R.probe(mh,nbr)
    print(name,'band probes: value mismatches',(v1!=v2).sum(),'active mismatches',(a1!=a2).sum(), 'n',len(nbr))
    R.free(mh)
t0=time.time(); g=ctx.build_sphere(100.0); ctx.synchronize(); print('build sphere s',time.time()-t0)
rs=R.sphere(100.0); compare('sphere100',g,rs)
fg=ctx.build_fog(g); rf=R.fog_from_levelset(rs); compare('fog100',fg,rf)
g2=ctx.build_sphere(5.0,(20,0,0),0.5,2.0); compare('sphere5',g2,R.sphere(5.0,(20,0,0),0.5,2.0),100000)
tq=ctx.build_torus(60.0,25.0); compare('torus60',tq,R.torus(60.0,25.0))
rng=np.random.default_rng(20240607)
S=np.column_stack([rng.uniform(-150,150,(24,3)),rng.uniform(10,40,24)])
u=ctx.build_spheres(S); compare('union24',u,R.spheres_union(S))
# C2 scale
t0=time.time(); big=ctx.build_torus(650.0,325.0); ctx.synchronize(); tb=time.time()-t0
i=big.info; print('C2 torus build s %.3f'%tb,'bytes',i.bytes,'leaves',i.leaf_count,'voxels',i.active_voxels,'lower',i.lower_count,'upper',i.upper_count,'tiles',i.root_tiles,'bbox',list(i.node_bbox))
W,H=1920,1080
R_,r_=650.0,325.0
cam=api.vdb_render_camera(W,H,(0,1.5*R_,3*(R_+r_)),(0,0,0))
for shk in (abi.SHADER_DIFFUSE, abi.SHADER_NORMAL):
    sh=api.make_shader(shk)
    film=api.PinnedArray((H,W,4),np.float32); film.array[...]=(0,0,0,1)
    for it in range(3):
        t0=time.time(); ctx.render_levelset(big,cam,sh,film.array); te=time.time()-t0
        ms,_=ctx.last_kernel_ms(); print('C2 shader',shk,'e2e ms %.3f kernel ms %.3f'%(te*1e3,ms),'hits',(film.array[...,:3].sum(axis=2)>0).sum())
c=ctx.count_levelset(big,cam); d=c.as_dict(); print('C2 counters',d)
n=d['rays']; B=32*d['root_probes']/n+16*d['upper_probes']/n+16*d['lower_probes']/n+12*d['voxel_probes']/n+32*d['stencil_refills']/n+16
print('C2 B/ray',B)
tf0=time.time(); bfog=ctx.build_fog(big) if False else None
