"""x-slab partition of the channel over GPUs: the host-side rules that replace the reference's
MPI_Cart_create 2-D decomposition (include/LBMGrid.h:347-392) and its neighbour discovery.

One process per GPU; rank r owns the contiguous columns [r*nx/world, (r+1)*nx/world) over the
full height (py == 1: no corner ghosts to exchange, so the N-slab run equals the 1-rank run bit
for bit, unlike the reference's own 2-D decompositions -- SURVEY.md F8).  After every collision
each interior slab face carries exactly the three populations that move across it:
EAST_GOING to the east neighbour's west ghost column, WEST_GOING the other way.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

EAST_GOING = (1, 5, 8)  # c_x = +1  (include/LBMConfig.h:13-25)
WEST_GOING = (3, 6, 7)  # c_x = -1
HALO_POPULATIONS = 3


@dataclass(frozen=True)
class Slab:
    rank: int
    world: int
    nx: int  # global
    ny: int
    periodic_x: bool = False

    def __post_init__(self):
        if self.world < 1 or not (0 <= self.rank < self.world):
            raise ValueError("bad rank/world %d/%d" % (self.rank, self.world))
        if self.nx % self.world:
            raise ValueError("nx=%d is not divisible by %d slabs" % (self.nx, self.world))  # as include/LBMGrid.h:358

    @property
    def lnx(self) -> int:
        return self.nx // self.world

    @property
    def x_start(self) -> int:
        return self.rank * self.lnx

    @property
    def west(self) -> int:
        """Neighbour rank on the -x side, or -1 at the inlet."""
        if self.world == 1:
            return -1
        if self.rank > 0:
            return self.rank - 1
        return self.world - 1 if self.periodic_x else -1

    @property
    def east(self) -> int:
        if self.world == 1:
            return -1
        if self.rank < self.world - 1:
            return self.rank + 1
        return 0 if self.periodic_x else -1

    @property
    def has_inlet(self) -> bool:
        return self.rank == 0 and not self.periodic_x

    @property
    def has_outlet(self) -> bool:
        return self.rank == self.world - 1 and not self.periodic_x

    def halo_bytes_per_step(self) -> int:
        """fp64 bytes this slab SENDS per iteration (the reference sends 9*ny per face)."""
        faces = (self.west >= 0) + (self.east >= 0)
        return faces * HALO_POPULATIONS * self.ny * 8

    def owner_of_column(self, x: int) -> int:
        return x // self.lnx


def env_rank_world():
    """(rank, world, local_rank) as torchrun exports them; (0, 1, 0) when started directly."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0")) if world > 1 else 0
    local = int(os.environ.get("LOCAL_RANK", str(rank))) if world > 1 else 0
    return rank, world, local


def create_slab_solver(params, dist=None, device=None):
    """Solver for this process's slab.  With torch.distributed initialised (`dist`), rank 0's NCCL
    unique id is broadcast to the other ranks; the halo traffic itself never touches torch."""
    from . import binding

    rank, world, local = env_rank_world()
    slab = Slab(rank, world, params.nx, params.ny, bool(params.flags & binding.FLAG_PERIODIC_X))
    nccl_id = None
    if world > 1:
        if dist is None:
            raise RuntimeError("world > 1 needs an initialised torch.distributed to hand the NCCL id around")
        box = [binding.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]
    s = binding.Solver(params, device=local if device is None else device, rank=rank, world=world, nccl_id=nccl_id)
    return s, slab
