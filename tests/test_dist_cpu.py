"""world_size-2 gloo test (CPU) of the N>1 host logic: row shards, global-pixel-keyed RNG, band gather, max-over-ranks."""
import os
import subprocess
import sys

import pytest

from chrono_photo_b200.sharding import shard_rows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_partition():
    for h in (1, 7, 37, 4000, 2160):
        for world in (1, 2, 3, 4, 8):
            if world > h:
                continue
            spans = [shard_rows(h, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(s[1] for s in spans) == h
            for a, b in zip(spans, spans[1:]):
                assert a[0] + a[1] == b[0]
            assert max(s[1] for s in spans) - min(s[1] for s in spans) <= 1
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


@pytest.mark.timeout(300)
def test_two_rank_gloo_row_shard_and_gather():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dist_worker.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=280)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "DIST_OK 2" in p.stdout
