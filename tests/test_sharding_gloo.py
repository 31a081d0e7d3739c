"""N > 1 host-side logic on CPU: frame sharding + result gather over a world_size-2 gloo
group (the compute function is the oracle port; on the GPU box bench.py plugs in the CUDA
matcher through the same code)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_frames_for_rank_partition():
    from tea_stereo_matching_b200.sharding import frames_for_rank

    for n in (0, 1, 7, 64):
        for world in (1, 2, 3, 8):
            shards = [frames_for_rank(n, world, r) for r in range(world)]
            assert sorted(sum(shards, [])) == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        frames_for_rank(4, 2, 2)


WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["TSM_ROOT"])
import oracle
from tea_stereo_matching_b200.sharding import run_sharded, max_over_ranks, sum_over_ranks
from tea_stereo_matching_b200.synth import synth_v1
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
port = oracle.Port()
pairs = [synth_v1(40, 56, 8, seed=100 + i) for i in range(5)]
res = run_sharded(pairs, lambda l, r: port.compute(l, r, 8), dist=dist, gather=True)
assert max_over_ranks(float(rank), dist) == world - 1
assert sum_over_ranks(1.0, dist) == world
if rank == 0:
    single = [port.compute(l, r, 8) for l, r in pairs]
    assert len(res) == 5 and all(np.array_equal(a, b) for a, b in zip(res, single))
    print("SHARDING_OK")
else:
    assert res is None
dist.destroy_process_group()
"""


def test_two_rank_gloo_shard_and_gather(port, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, TSM_ROOT=str(ROOT), OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29517", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDING_OK" in r.stdout
