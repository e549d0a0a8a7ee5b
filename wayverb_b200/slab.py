"""z-slab bookkeeping shared by bench.py, the tests and the docs.

The C++ side (csrc/wg_host.cu: create_impl, local_offset, exchange_ghosts)
implements exactly this plan; it is restated here so that the N>1 protocol can
be exercised on CPU (gloo, world_size 2) without a GPU.

Layout of one rank's slab: local planes 0 .. nzl+1, plane 0 and nzl+1 are ghost
copies of the z-neighbours' edge planes (or stay zero at the mesh ends).
After every step each rank sends its first and last OWNED plane of the array
that was just written to the neighbour below / above and receives their edge
planes into its ghost planes -- one message per face per step.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

from .waveguide import slab_range


@dataclass(frozen=True)
class Transfer:
    peer: int
    send_plane: int  # local plane index sent to `peer`
    recv_plane: int  # local plane index filled from `peer`


@dataclass(frozen=True)
class SlabPlan:
    rank: int
    nranks: int
    dz: int
    z_begin: int
    z_end: int

    @property
    def nzl(self) -> int:
        return self.z_end - self.z_begin

    @property
    def node_planes(self) -> Tuple[int, int]:
        """global planes whose condensed nodes this rank needs (owned + ghosts in the mesh)"""
        return max(self.z_begin - 1, 0), min(self.z_end + 1, self.dz)

    def transfers(self) -> List[Transfer]:
        t = []
        if self.rank > 0:
            t.append(Transfer(self.rank - 1, 1, 0))
        if self.rank < self.nranks - 1:
            t.append(Transfer(self.rank + 1, self.nzl, self.nzl + 1))
        return t

    def local_plane(self, z: int):
        """local plane holding a copy of global plane z (owned or ghost), else None"""
        lz = z - self.z_begin + 1
        return lz if 0 <= lz <= self.nzl + 1 and 0 <= z < self.dz else None

    def owns(self, z: int) -> bool:
        return self.z_begin <= z < self.z_end


def make_plan(dz: int, rank: int, nranks: int) -> SlabPlan:
    z0, z1 = slab_range(dz, rank, nranks)
    return SlabPlan(rank, nranks, dz, z0, z1)


def ray_range(total_rays: int, rank: int, nranks: int) -> Tuple[int, int]:
    """The ray path's sharding: rays are independent, so rank r traces the
    global rays [begin, end) -- `ray_index_base = begin`, `total_rays` = all ranks' rays --
    against a replicated scene, and the histograms are summed over ranks
    (wvb_rt_allreduce_histogram; sum_histograms, stochastic/postprocessing.h:72-90)."""
    return total_rays * rank // nranks, total_rays * (rank + 1) // nranks
