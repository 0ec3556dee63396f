import os,sys
sys.path.insert(0,'/root/repo')
import torch
from pgr_raytracing_project_b200 import scenes
from pgr_raytracing_project_b200.context import RenderContext
W,H=1920,1080
s=scenes.random_triangles(1_000_000); ctx=RenderContext(0); ctx.set_scene(s)
ctx.set_camera(s.camera.position, s.camera.target, s.camera.up, s.camera.fov)
host=torch.empty((H,W,3),dtype=torch.float32,pin_memory=True)
for k in range(6):
    print("--- frame",k,file=sys.stderr,flush=True)
    ctx.render_host(W,H,1,1,seed=1,out=host.numpy())
