"""Cascade-parallel sharding across the GPUs of one box.

Cascades are independent (every reference dispatch addresses exactly one layer through its
``cascade_index`` push constant, assets/water/wave_generator.gd:65-85,96-97), so the path shards with NO
data-path collective: rank r owns cascades {i : i mod R == r}, keeps their spectrum, scratch and foam
layers resident, and never exchanges anything inside an update.  torch.distributed is used for the
plumbing only: an optional parameter broadcast (<= 200 B per cascade) and an optional all-gather of
the finished RGBA16F layers when one consumer wants the whole Texture2DArray.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np


def owned_cascades(num_cascades: int, rank: int, world: int) -> List[int]:
    """Round-robin partition: global cascade indices owned by `rank` (ascending)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"rank {rank} / world {world}")
    return list(range(rank, num_cascades, world))


def owner_of(cascade: int, world: int) -> int:
    return cascade % world


class ShardedWaveGenerator:
    """A WaveGenerator per rank over the cascades that rank owns.

    ``generator_factory(map_size, num_local_cascades)`` must return an object with the WaveGenerator
    interface (update, update_all, _process, maps_to_host, free); the default builds the CUDA generator on
    ``device``.  Tests inject the CPU oracle here to exercise the multi-process logic under gloo."""

    def __init__(self, map_size: int, rank: Optional[int] = None, world: Optional[int] = None, device: int = 0,
                 generator_factory: Optional[Callable] = None, group=None):
        self.map_size = int(map_size)
        self.group = group
        if rank is None or world is None:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(group), dist.get_world_size(group)
            else:
                rank, world = 0, 1
        self.rank, self.world = int(rank), int(world)
        self.device = device
        self._factory = generator_factory or self._cuda_factory
        self.gen = None
        self.owned: List[int] = []
        self.num_cascades = 0

    def _cuda_factory(self, map_size: int, num_local: int):
        from .wave_generator import WaveGenerator
        g = WaveGenerator(device=self.device)
        g.map_size = map_size
        g.init_gpu(max(1, num_local))
        return g

    # -- parameters ------------------------------------------------------------------------------------
    def broadcast_parameters(self, parameters, src: int = 0):
        """Replicates rank `src`'s parameter list on every rank (objects are pickled; tiny)."""
        if self.world == 1:
            return parameters
        import torch.distributed as dist
        box = [parameters if self.rank == src else None]
        dist.broadcast_object_list(box, src=src, group=self.group)
        return box[0]

    def _local(self, parameters: Sequence):
        n = len(parameters)
        if self.gen is None or n != self.num_cascades:
            if self.gen is not None:
                self.gen.free()
            self.num_cascades = n
            self.owned = owned_cascades(n, self.rank, self.world)
            self.gen = self._factory(self.map_size, len(self.owned))
            for i in self.owned:            # a new generator starts with empty spectra (water.gd:84-87 marks them all dirty)
                parameters[i].should_generate_spectrum = True
        return [parameters[i] for i in self.owned]

    # -- WaveGenerator surface (wave_generator.gd:56-63,90-109), restricted to the owned cascades ----------
    def update(self, delta: float, parameters: Sequence) -> None:
        local = self._local(parameters)
        if local:
            self.gen.update(delta, local)

    def update_all(self, delta: float, parameters: Sequence) -> None:
        local = self._local(parameters)
        if local:
            self.gen.update_all(delta, local)

    def _process(self, delta: float = 0.0) -> None:
        if self.gen is not None and self.owned:
            self.gen._process(delta)

    def local_maps_to_host(self):
        """(displacement, normal) float16 [len(owned), N, N, 4] of the owned cascades."""
        N = self.map_size
        if not self.owned:
            return np.zeros((0, N, N, 4), np.float16), np.zeros((0, N, N, 4), np.float16)
        return self.gen.maps_to_host(0, len(self.owned))

    def gather_maps(self):
        """All cascades' maps on every rank, in global cascade order: ([C,N,N,4], [C,N,N,4]) float16.
        One all_gather of the (padded) local layers per map -- NOT part of the timed throughput path."""
        d, n = self.local_maps_to_host()
        if self.world == 1:
            return d, n
        import torch
        import torch.distributed as dist
        N, C = self.map_size, self.num_cascades
        per = (C + self.world - 1) // self.world                       # padded layers per rank
        backend = dist.get_backend(self.group)
        dev = torch.device("cuda", self.device) if backend == "nccl" else torch.device("cpu")
        out = []
        for arr in (d, n):
            local = torch.zeros((per, N, N, 4), dtype=torch.float16, device=dev)
            if len(self.owned):
                local[:len(self.owned)] = torch.from_numpy(arr).to(dev)
            full = torch.empty((self.world * per, N, N, 4), dtype=torch.float16, device=dev)
            dist.all_gather_into_tensor(full, local, group=self.group)
            full = full.cpu().numpy().reshape(self.world, per, N, N, 4)
            res = np.empty((C, N, N, 4), np.float16)
            for r in range(self.world):
                own = owned_cascades(C, r, self.world)
                res[own] = full[r, :len(own)]
            out.append(res)
        return out[0], out[1]

    def free(self) -> None:
        if self.gen is not None:
            self.gen.free()
            self.gen = None
