"""Multi-GPU tests proper (need >= 2 CUDA devices; skipped otherwise): one process per GPU over NCCL, every
partition mode of pgr_raytracing_project_b200.multigpu against the single-GPU frame.  `peer` / `peer_samples`
exercise the CUDA-IPC frame sharing: ranks > 0 render straight into rank 0's memory over NVLink."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pgr_raytracing_project_b200 import scenes
    from pgr_raytracing_project_b200.context import RenderContext
    from pgr_raytracing_project_b200.multigpu import DistributedRenderer
    ctx = RenderContext(rank)
    W, H = 320, 200
    for name, scene, spp, depth in [("tris_d1", scenes.random_triangles(20_000, seed=8), 2 * world, 1),
                                    ("default_d4", scenes.default_scene(), 2 * world, 4)]:
        ctx.set_scene(scene)
        c = scene.camera
        ctx.set_camera(c.position, c.target, c.up, c.fov)
        full = ctx.render(W, H, spp, depth, seed=5).cpu().numpy() if rank == 0 else None
        for mode in ("tiles", "samples", "peer", "peer_samples", "peer+nccl", "peer_samples+nccl"):
            r = DistributedRenderer(ctx, rank, world, mode=mode.split("+")[0], barrier="nccl" if mode.endswith("+nccl") else "flag")
            for slot in (0, 1, 0):                                  # buffer sets are reusable
                frame = r.render(W, H, spp, depth, seed=5, slot=slot)
            torch.cuda.synchronize()
            if rank == 0:
                np.save(os.path.join(out_dir, f"{name}_{mode}.npy"), frame.cpu().numpy())
            dist.barrier()
            # a pipeline of DIFFERENT consecutive frames with no host synchronisation in between (the renderer alternates
            # its two shared buffers itself): every frame is consumed by a copy enqueued on the stream before the next
            # render call -- what the lifetime rule of DistributedRenderer.render asks for -- while the other ranks run ahead
            kept = []
            for k in range(6):
                frame = r.render(W, H, spp, depth, seed=5, sample_offset=k * spp)
                if rank == 0:
                    kept.append(frame.clone())
            torch.cuda.synchronize()
            if rank == 0:
                np.save(os.path.join(out_dir, f"{name}_{mode}_seq.npy"), torch.stack(kept).cpu().numpy())
            dist.barrier()
            r.close()
            if mode == "peer":                                      # frames assembled in shared HOST memory, every GPU pushing its own tiles
                hosts = []
                for k in range(5):
                    hf = r.render_host(W, H, spp, depth, seed=5, sample_offset=k * spp)
                    if rank == 0:
                        hosts.append(hf.copy())
                torch.cuda.synchronize()
                if rank == 0:
                    np.save(os.path.join(out_dir, f"{name}_host_seq.npy"), np.stack(hosts))
                dist.barrier()
            if mode == "peer":                                      # a closed renderer starts afresh (frames, sync words, counters)
                frame = r.render(W, H, spp, depth, seed=5)
                torch.cuda.synchronize()
                if rank == 0:
                    np.save(os.path.join(out_dir, f"{name}_reopened.npy"), frame.cpu().numpy())
                dist.barrier()
                r.close()
        if rank == 0:
            np.save(os.path.join(out_dir, f"{name}_full.npy"), full)
            seq = [ctx.render(W, H, spp, depth, seed=5, sample_offset=k * spp).cpu().numpy() for k in range(6)]
            np.save(os.path.join(out_dir, f"{name}_full_seq.npy"), np.stack(seq))
    ctx.close()
    dist.destroy_process_group()


def test_all_partition_modes_match_single_gpu(tmp_path):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 CUDA devices")
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for name in ("tris_d1", "default_d4"):
        full = np.load(tmp_path / f"{name}_full.npy")
        seq = np.load(tmp_path / f"{name}_full_seq.npy")
        assert not np.array_equal(seq[0], seq[1])                   # consecutive frames really differ
        for mode in ("tiles", "peer", "peer+nccl"):
            assert np.array_equal(np.load(tmp_path / f"{name}_{mode}.npy"), full), (name, mode)   # bit-identical
            assert np.array_equal(np.load(tmp_path / f"{name}_{mode}_seq.npy"), seq), (name, mode)
        assert np.array_equal(np.load(tmp_path / f"{name}_reopened.npy"), full), name
        assert np.array_equal(np.load(tmp_path / f"{name}_host_seq.npy"), seq[:5]), name
        for mode in ("samples", "peer_samples", "peer_samples+nccl"):
            np.testing.assert_allclose(np.load(tmp_path / f"{name}_{mode}.npy"), full, atol=3e-6)  # re-associated sum
            np.testing.assert_allclose(np.load(tmp_path / f"{name}_{mode}_seq.npy"), seq, atol=3e-6)
