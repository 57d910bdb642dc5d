"""CPU test of the N>1 host logic with world_size 2 over gloo: chunk partition, ghost / send lists, halo exchange
pattern and rank-ordered reductions — emulating the device path in numpy and checking against the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, outq):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import oracle_ffi as o
    from kryst_b200 import parallel
    try:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        # the communicator id travels through the process group exactly as in bench.py
        uid = parallel.broadcast_unique_id(lambda: bytes(range(128)), rank)
        assert uid == bytes(range(128))
        for kind, N in (("poisson3d", 8), ("convdiff2d", 12), ("varcoef27", 5)):
            n, lo, hi, rp, ci, v = parallel.shard_stencil(kind, N, world, rank)
            Ao = o.stencil(kind, N)
            ghosts, local, recv_from = parallel.host_ghost_plan(n, world, rank, rp, ci)
            assert np.array_equal(ghosts, o.ghost_list(Ao, lo, hi))
            allg = [None] * world
            dist.all_gather_object(allg, ghosts)
            sends = parallel.send_lists(allg, n, world, rank)
            # halo exchange of x through the process group
            xg = np.random.default_rng(3).standard_normal(n)
            x_local = np.concatenate([xg[lo:hi], np.zeros(ghosts.size)])
            packed = {q: x_local[rows - lo] for q, rows in sends.items()}
            allp = [None] * world
            dist.all_gather_object(allp, packed)
            for q, (off, cnt) in recv_from.items():
                x_local[(hi - lo) + off:(hi - lo) + off + cnt] = allp[q][rank]
            # local SpMV in stored (== ascending global column) order
            y = np.zeros(hi - lo)
            for i in range(hi - lo):
                s = 0.0
                for p in range(int(rp[i]), int(rp[i + 1])):
                    s = s + v[p] * x_local[local[p]]
                y[i] = s
            assert np.array_equal(y, o.spmv(Ao, xg)[lo:hi])
            # rank-ordered reduction == the oracle's sharded dot
            parts = [None] * world
            dist.all_gather_object(parts, o.dot(xg[lo:hi], xg[lo:hi]))
            s = parts[0]
            for r in range(1, world):
                s = s + parts[r]
            assert s == o.dot(xg, xg, nshards=world)
        dist.barrier()
        dist.destroy_process_group()
        outq.put((rank, "ok"))
    except Exception as e:   # pragma: no cover
        import traceback
        outq.put((rank, traceback.format_exc()))


def test_two_rank_host_logic_gloo(built):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29533
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
