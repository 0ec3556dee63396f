"""Option-combination check (also the driver for compute-sanitizer where that is available): a few small launches of the round-2 kernels (k_tiny both modes, shared host frames, treelet, fold)."""
import mmap, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pgr_raytracing_project_b200 import scenes
from pgr_raytracing_project_b200.context import RenderContext
ctx = RenderContext(0)
for make in (scenes.cornell_box, scenes.default_scene):
    s = make(); W, H = 72, 40
    ctx.set_scene(s); ctx.set_camera_array(s.camera.as_array(W / H))
    ctx.set_option("kernel", 1); ref = ctx.render(W, H, 5, 4, seed=3).clone()
    ctx.set_option("kernel", 5)
    for mode in (0, 1):
        ctx.set_option("tiny_mode", mode)
        for th in (256, 128):
            ctx.set_option("tiny_threads", th)
            assert torch.equal(ctx.render(W, H, 5, 4, seed=3), ref), (s.name, mode, th)
    ctx.set_option("tiny_mode", 0); ctx.set_option("tiny_threads", 256); ctx.set_option("kernel", -1)
s = scenes.random_triangles(5000, seed=3); W, H = 96, 64
ctx.set_scene(s); ctx.set_camera_array(s.camera.as_array(W / H))
ref1 = ctx.render(W, H, 1, 1, seed=4).clone(); ref3 = ctx.render(W, H, 3, 1, seed=4).clone(); refd = ctx.render(W, H, 2, 3, seed=4).clone()
for opt, val in (("treelet", 4), ("fold", 1)):
    ctx.set_option(opt, val)
    assert torch.equal(ctx.render(W, H, 1, 1, seed=4), ref1) and torch.equal(ctx.render(W, H, 3, 1, seed=4), ref3) and torch.equal(ctx.render(W, H, 2, 3, seed=4), refd), opt
    ctx.set_option(opt, 0)
fb = W * H * 12; off = (fb + 4095) & ~4095
mm = mmap.mmap(-1, off + 4096); buf = np.frombuffer(mm, dtype=np.uint8); addr = buf.ctypes.data
alias = ctx.host_register(addr, off + 4096)
for rank in range(2):
    ctx.render_tiles_host(W, H, rank, 2, 3, 1, 4, 0, alias, alias + off + 4 * rank, 1)
ctx.host_wait(addr + off, 2, 1)
assert np.array_equal(buf[:fb].view(np.float32).reshape(H, W, 3), ref3.cpu().numpy())
torch.cuda.synchronize()
ctx.host_unregister(addr)
print("sanitizer driver ok")
