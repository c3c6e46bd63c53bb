"""world_size-2 gloo run of the sequence-sharding host logic (unicorn_b200/parallel.py) on CPU."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from unicorn_b200.parallel import run_sharded, shard_sequences
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seqs = [list(range(3 + i)) for i in range(5)]  # 5 sequences of 3..7 frames
    res = run_sharded(seqs, lambda i, s: (len(s), 1))
    q.put((rank, res["frames"], res["tracks"], res["shard"], shard_sequences(5, rank, world)))
    dist.destroy_process_group()


def test_two_ranks_cover_all_sequences_once():
    from unicorn_b200.parallel import shard_sequences
    for n in (1, 2, 5, 8, 17):
        for w in (1, 2, 4, 8):
            owned = sorted(i for r in range(w) for i in shard_sequences(n, r, w))
            assert owned == list(range(n))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(o[1] == 3 + 4 + 5 + 6 + 7 and o[2] == 5 for o in out)  # every rank sees the global totals
    assert sorted(out[0][3] + out[1][3]) == [0, 1, 2, 3, 4]
