"""Worker for the multi-GPU parity test: torch.distributed.run, one rank per GPU, NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import godotoceanwaves_b200 as gow  # noqa: E402
from conftest import demo_params  # noqa: E402
from godotoceanwaves_b200.sharding import ShardedWaveGenerator  # noqa: E402


def main():
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    N, C = 256, 8
    params = [demo_params(gow.WaveCascadeParameters, c) for c in range(C)]
    sh = ShardedWaveGenerator(N, device=local_rank)
    for _ in range(3):
        sh.update_all(0.02, params)
    d, n = sh.gather_maps()
    # single-GPU reference of all cascades on this rank's own GPU
    ref = gow.WaveGenerator(device=local_rank)
    ref.map_size = N
    ref.init_gpu(C)
    rp = [demo_params(gow.WaveCascadeParameters, c) for c in range(C)]
    for _ in range(3):
        ref.update_all(0.02, rp)
    rd, rn = ref.maps_to_host()
    assert np.array_equal(d.view(np.uint16), rd.view(np.uint16)), "sharded != single GPU (displacement)"
    assert np.array_equal(n.view(np.uint16), rn.view(np.uint16)), "sharded != single GPU (normal/foam)"
    dist.barrier()
    if rank == 0:
        print(f"SHARDING_GPU_OK world={world}")
    sh.free(); ref.free()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
