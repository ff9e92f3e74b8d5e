"""CPU, world_size 2, gloo: the only multi-process logic of the path -- contiguous sharding of the
image list and the all-gather of container byte counts (l3c_pytorch_b200/dist.py)."""
import os
import socket

import torch
import torch.multiprocessing as mp

from l3c_pytorch_b200 import dist as l3c_dist


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 16, 128, 129):
        for world in (1, 2, 3, 8):
            spans = [l3c_dist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_items, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    r, w, _ = l3c_dist.init_from_env(backend='gloo')
    lo, hi = l3c_dist.shard_bounds(n_items, r, w)
    local = [1000 + 7 * i for i in range(lo, hi)]          # "container sizes" of my images
    counts = l3c_dist.gather_byte_counts(local, n_items, r, w, device=torch.device('cpu'))
    q.put((rank, counts, l3c_dist.global_bpsp(counts, 3 * 32 * 32)))
    torch.distributed.destroy_process_group()


def test_gather_byte_counts_world2():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    for n_items in (5, 4):
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
        for p in procs:
            p.start()
        results = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(60)
            assert p.exitcode == 0
        want = [1000 + 7 * i for i in range(n_items)]
        for _, counts, bpsp in results:
            assert counts == want
            assert abs(bpsp - sum(want) * 8 / (n_items * 3072)) < 1e-12
        port += 1
