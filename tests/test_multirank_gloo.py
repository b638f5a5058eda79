"""World-size-2 (and 3) CPU test of the multi-GPU host logic over gloo: every rank renders only the tiles it owns
(vdbrt_partition semantics, here evaluated by the oracle port because there is no GPU), the tiles are gathered with the
product's TileGather, and rank 0's frame must equal the unpartitioned render bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, buf_path, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openvdb_b200 import api, _abi as abi
    from openvdb_b200.frame import TileGather
    from tests import refapi
    oracle = refapi.Oracle()
    buf = refapi.aligned_copy(np.load(buf_path))
    og = oracle.open(buf)
    W, H, TW, TH = 128, 96, 32, 24
    cam = api.vdb_render_camera(W, H, (10.0, 20.0, 140.0), (0.0, 0.0, 0.0))
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    film = refapi.new_film(W, H, (0.25, 0.5, 0.75, 1.0))
    oracle.render_levelset(og, cam, sh, film, part=api.partition(rank, world, TW, TH))
    t = torch.from_numpy(film)
    TileGather(H, W, TH, TW, rank, world, "cpu").gather(t)
    if rank == 0:
        whole = refapi.new_film(W, H, (0.25, 0.5, 0.75, 1.0))
        oracle.render_levelset(og, cam, sh, whole)
        np.save(out_path, np.stack([t.numpy(), whole]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tile_partition_and_gather(tmp_path, world):
    from tests import refapi
    if not os.path.exists(refapi.REF_SO):
        pytest.skip("needs a grid from the reference generators")
    ref = refapi.Ref()
    buf = ref.nanovdb(ref.sphere(40.0, (3.0, -2.0, 1.0)))
    buf_path = str(tmp_path / "grid.npy")
    out_path = str(tmp_path / "out.npy")
    np.save(buf_path, buf)
    mp.spawn(_worker, args=(world, _free_port(), buf_path, out_path), nprocs=world, join=True)
    got, want = np.load(out_path)
    assert (want[..., 0] != 0.25).sum() > 1000
    assert np.array_equal(got, want)


def test_tile_bookkeeping():
    from openvdb_b200.frame import TileGather
    g = TileGather(1080, 1920, 60, 64, 0, 8, "cpu")
    assert g.ntiles == 540 and g.per_rank == 68
    owned = [set(g.owned(r)) for r in range(8)]
    assert set().union(*owned) == set(range(540)) and sum(len(o) for o in owned) == 540
    film = torch.arange(1080 * 1920 * 4, dtype=torch.float32).view(1080, 1920, 4)
    assert torch.equal(g.untile(g.tiles(film)), film)
    with pytest.raises(ValueError):
        TileGather(1080, 1920, 64, 64, 0, 2, "cpu")
