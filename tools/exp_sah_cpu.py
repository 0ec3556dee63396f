"""CPU study behind option builder 2: a binned-SAH tree (python prototype, leaves of <= 4 forced) against the reference's median
split on a triangle soup of C3's density -- node records and triangle tests per ray, counted by the oracle walking both trees
(interior random rays and primary rays); the hits must be identical.  usage: python tools/exp_sah_cpu.py [n_triangles]"""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as orc
from pgr_raytracing_project_b200 import scenes
sys.setrecursionlimit(100000)
N=int(sys.argv[1]) if len(sys.argv)>1 else 200000
ext=10.0*(N/1e6)**(1/3)
s=scenes.random_triangles(N, extent=ext)
V=s.vertices.reshape(-1,3,3).astype(np.float64)
lo=V.min(1); hi=V.max(1); cen=0.5*(lo+hi)
NODE=orc.NODE_DTYPE
def build(split):
    nodes=[]; order=[]
    nodes.append(None); nodes.append(None)   # root, pad
    def box(idx): return lo[idx].min(0), hi[idx].max(0)
    def area(a,b):
        e=b-a; return e[0]*e[1]+e[1]*e[2]+e[2]*e[0]
    def rec(idx, slot):
        bl,bh=box(idx)
        n=len(idx)
        if n<=4:
            nodes[slot]=(bl,len(order),bh,n); order.extend(idx.tolist()); return
        l,r=split(idx,bl,bh,area)
        a=len(nodes); nodes.append(None); nodes.append(None)
        nodes[slot]=(bl,a,bh,0)
        rec(l,a); rec(r,a+1)
    rec(np.arange(N),0)
    arr=np.zeros(len(nodes),dtype=NODE)
    pad=2.0**-16*np.abs(np.concatenate([lo,hi])).max()
    for k,nd in enumerate(nodes):
        if nd is None: continue
        arr[k]['bmin']=(nd[0]-pad).astype(np.float32); arr[k]['a']=nd[1]; arr[k]['bmax']=(nd[2]+pad).astype(np.float32); arr[k]['b']=nd[3]
    return arr, np.array(order,dtype=np.int32)
def median_split(idx,bl,bh,area):
    ax=int(np.argmax(bh-bl)); o=idx[np.argsort(cen[idx,ax],kind='stable')]; m=len(o)//2
    return o[:m],o[m:]
def sah_split(idx,bl,bh,area,B=16):
    best=None
    cl=cen[idx].min(0); ch=cen[idx].max(0)
    for ax in range(3):
        if ch[ax]<=cl[ax]: continue
        b=np.minimum(((cen[idx,ax]-cl[ax])/(ch[ax]-cl[ax])*B).astype(int),B-1)
        for k in range(1,B):
            L=idx[b<k]; R=idx[b>=k]
            if len(L)==0 or len(R)==0: continue
            c=area(lo[L].min(0),hi[L].max(0))*len(L)+area(lo[R].min(0),hi[R].max(0))*len(R)
            if best is None or c<best[0]: best=(c,L,R)
    if best is None: return median_split(idx,bl,bh,area)
    return best[1],best[2]
rng=np.random.default_rng(7)
R=20000
o=rng.uniform(-0.9*ext,0.9*ext,(R,3)).astype(np.float32); d=rng.normal(size=(R,3)); d/=np.linalg.norm(d,axis=1,keepdims=True); d=d.astype(np.float32)
O=orc.OracleScene(s)
cam=s.camera.as_array(16/9); O.set_camera(cam)
res={}
for name,sp in (("median",median_split),("sah16",sah_split)):
    t0=time.time(); arr,order=build(sp); tb=time.time()-t0
    O.set_bvh(arr,order)
    prim,t,st=O.trace_rays(o,d)
    p2,t2,st2=O.trace_primary(480,270)
    res[name]=(prim,t)
    print(f"{name:8s} build {tb:6.1f}s nodes {len(arr)}  interior rays: nodes/ray {st[1]/R:6.1f} tris/ray {st[2]/R:5.1f}   primary: nodes/ray {st2[1]/st2[0]:6.1f} tris/ray {st2[2]/st2[0]:5.1f}",flush=True)
print("same hits:", np.array_equal(res['median'][0],res['sah16'][0]), np.array_equal(res['median'][1],res['sah16'][1]))
