"""The multi-GPU frame assembly without a collective: two processes (both on cuda:0 here; one per GPU in bench.py) render the
tiles they own straight into rank 0's film through a CUDA IPC mapping.  The assembled frame must equal a one-process render."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, TW, TH = 320, 240, 32, 24


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openvdb_b200 import api, _abi as abi
    ctx = api.Context(0)
    grid = ctx.build_sphere(60.0, (5.0, -3.0, 2.0))
    cam = api.vdb_render_camera(W, H, (20.0, 30.0, 200.0), (0.0, 0.0, 0.0))
    sh = api.make_shader(abi.SHADER_NORMAL)

    def exchange(h):
        t = torch.from_numpy(h.copy())
        dist.broadcast(t, src=0)
        return t.numpy()

    shared = api.SharedFilm(ctx, H, W, rank, exchange)
    opts = ctx.ls_opts(part=api.partition(rank, world, TW, TH), uniform_bg=True)
    ctx.render_levelset(grid, cam, sh, shared.ptr, width=W, height=H, memspace=abi.MEM_DEVICE, bg=(0.2, 0.3, 0.4, 1.0), opts=opts)
    ctx.synchronize()
    dist.barrier()                      # every rank's tiles are in rank 0's film
    if rank == 0:
        got = np.empty((H, W, 4), np.float32)
        api.memcpy(ctx, got.ctypes.data, shared.ptr, got.nbytes, 1)
        ctx.synchronize()
        want = np.empty((H, W, 4), np.float32)
        want[...] = (0.2, 0.3, 0.4, 1.0)
        ctx.render_levelset(grid, cam, sh, want)
        np.save(out_path, np.stack([got, want]))
    dist.barrier()
    shared.close()
    ctx.close()
    dist.destroy_process_group()


def _host_worker(rank, world, port, out_path):
    """the end-to-end form: ONE host film in shared memory, page-locked by both processes, written in place by their kernels"""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openvdb_b200 import api, _abi as abi
    ctx = api.Context(0)
    grid = ctx.build_sphere(60.0, (5.0, -3.0, 2.0))
    cam = api.vdb_render_camera(W, H, (20.0, 30.0, 200.0), (0.0, 0.0, 0.0))
    sh = api.make_shader(abi.SHADER_DIFFUSE)

    def exchange(name):
        box = [name]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    shared = api.SharedHostFilm(H, W, rank, exchange)
    rng = np.random.default_rng(3)
    old = rng.random((H, W, 4)).astype(np.float32)          # an arbitrary old film: misses must keep it
    if rank == 0:
        shared.array[...] = old
    dist.barrier()
    ctx.render_levelset(grid, cam, sh, shared.array, opts=ctx.ls_opts(part=api.partition(rank, world, TW, TH)))
    dist.barrier()                      # every rank's pixels are in the host film
    if rank == 0:
        want = old.copy()
        ctx.render_levelset(grid, cam, sh, want)
        np.save(out_path, np.stack([np.array(shared.array), want]))
    dist.barrier()
    shared.close()
    ctx.close()
    dist.destroy_process_group()


def test_shared_host_film_equals_single_process_render(tmp_path):
    out_path = str(tmp_path / "out.npy")
    mp.spawn(_host_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    got, want = np.load(out_path)
    assert np.array_equal(got, want)


def test_peer_written_frame_equals_single_process_render(tmp_path):
    out_path = str(tmp_path / "out.npy")
    mp.spawn(_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    got, want = np.load(out_path)
    assert (want[..., :3] != np.float32([0.2, 0.3, 0.4])).any(axis=2).sum() > 5000
    assert np.array_equal(got, want)
