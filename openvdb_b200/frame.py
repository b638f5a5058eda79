"""Multi-GPU frame assembly: the only exchange on this path.

The film is cut into tile_h x tile_w tiles numbered row-major; rank r renders tiles r, r+world, ... (vdbrt_partition) and
the finished tiles are gathered to rank 0 with ONE collective (torch.distributed gather: NCCL over NVLink on GPUs, gloo
in the CPU tests).  Pure tensor plumbing -- no pixel is computed here.
"""
import torch
import torch.distributed as dist


class TileGather:
    def __init__(self, height, width, tile_h, tile_w, rank, world, device, channels=4, dtype=torch.float32):
        if height % tile_h or width % tile_w:
            raise ValueError("film %dx%d is not a multiple of the %dx%d tile" % (width, height, tile_w, tile_h))
        self.H, self.W, self.th, self.tw, self.C = height, width, tile_h, tile_w, channels
        self.rank, self.world = rank, world
        self.ty, self.tx = height // tile_h, width // tile_w
        self.ntiles = self.ty * self.tx
        self.per_rank = (self.ntiles + world - 1) // world
        self.send = torch.zeros((self.per_rank, tile_h, tile_w, channels), dtype=dtype, device=device)
        self.recv = [torch.zeros_like(self.send) for _ in range(world)] if rank == 0 else None

    def owned(self, rank):
        """flat tile ids rendered by `rank`"""
        return range(rank, self.ntiles, self.world)

    def tiles(self, film):
        """[ntiles, th, tw, C] copy of the film in tile order"""
        return film.view(self.ty, self.th, self.tx, self.tw, self.C).permute(0, 2, 1, 3, 4).reshape(self.ntiles, self.th, self.tw, self.C)

    def untile(self, tiles):
        return tiles.view(self.ty, self.tx, self.th, self.tw, self.C).permute(0, 2, 1, 3, 4).reshape(self.H, self.W, self.C)

    def gather(self, film):
        """in place on rank 0: after the call rank 0's film holds every rank's tiles"""
        mine = self.tiles(film)[self.rank::self.world]
        self.send[:mine.shape[0]].copy_(mine)
        dist.gather(self.send, self.recv, dst=0)
        if self.rank == 0:
            tv = self.tiles(film)
            for q in range(1, self.world):
                n = tv[q::self.world].shape[0]
                tv[q::self.world] = self.recv[q][:n]
            film.copy_(self.untile(tv))
        return film
