"""Multi-process host logic of the cascade-parallel split on CPU: world_size 2, gloo."""
import os
import subprocess
import sys

import pytest

from godotoceanwaves_b200.sharding import owned_cascades, owner_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_is_a_disjoint_cover():
    for C in (1, 4, 5, 8, 128):
        for R in (1, 2, 4, 8):
            seen = []
            for r in range(R):
                own = owned_cascades(C, r, R)
                assert all(owner_of(i, R) == r for i in own) and own == sorted(own)
                seen += own
            assert sorted(seen) == list(range(C))
    # cfg4: 8 cascades -> 8/4/2/1 per GPU at R = 1/2/4/8
    assert [len(owned_cascades(8, 0, R)) for R in (1, 2, 4, 8)] == [8, 4, 2, 1]
    with pytest.raises(ValueError):
        owned_cascades(4, 2, 2)


def test_world_size_2_gloo_sharded_equals_single_process():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "_sharding_worker.py")]
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert res.returncode == 0 and "SHARDING_OK" in res.stdout, res.stdout[-3000:]
