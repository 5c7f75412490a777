"""Worker for tests/test_sharding_cpu.py: launched by torch.distributed.run with world_size 2 (gloo).
The CPU oracle is injected as the per-rank generator so that the multi-process host logic (partition,
parameter broadcast, per-rank update, gather) runs without a GPU."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import demo_params  # noqa: E402
from godotoceanwaves_b200.sharding import ShardedWaveGenerator, owned_cascades  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


class OracleBackend:
    """WaveGenerator-shaped wrapper of the oracle (tests only)."""

    def __init__(self, map_size, num_local):
        self.o = po.OracleWaveGenerator(map_size)
        self.o.init_gpu(max(1, num_local))

    def update(self, delta, params):
        self.o.update(delta, params)

    def update_all(self, delta, params):
        self.o.update_all(delta, params)

    def _process(self, delta=0.0):
        self.o.process()

    def maps_to_host(self, first, count):
        return (self.o.displacement_half()[first:first + count].copy(), self.o.normal_half()[first:first + count].copy())

    def free(self):
        pass


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    N, C = 128, 5
    # only rank 0 knows the parameters; everyone else receives them
    params = [demo_params(po.CascadeParams, c) for c in range(C)] if rank == 0 else None
    sh = ShardedWaveGenerator(N, generator_factory=lambda n, k: OracleBackend(n, k))
    params = sh.broadcast_parameters(params, src=0)
    assert len(params) == C and sh.world == world == 2
    for _ in range(2):
        sh.update_all(0.02, params)
    assert sh.owned == owned_cascades(C, rank, world)
    d, n = sh.gather_maps()
    # single-process reference on every rank
    ref = po.OracleWaveGenerator(N)
    rp = [demo_params(po.CascadeParams, c) for c in range(C)]
    for _ in range(2):
        ref.update_all(0.02, rp)
    assert np.array_equal(d.view(np.uint16), ref.displacement_map[:C]), "sharded displacement != single process"
    assert np.array_equal(n.view(np.uint16), ref.normal_map[:C]), "sharded normal/foam != single process"
    # owned parameter objects advanced in time, the others untouched
    for i, p in enumerate(params):
        assert (p.time == rp[i].time) == (i in sh.owned)
    dist.barrier()
    if rank == 0:
        print("SHARDING_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
