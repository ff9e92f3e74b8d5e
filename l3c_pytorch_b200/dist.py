"""Multi-GPU plumbing: one process per GPU, images (or crops) sharded in contiguous blocks, and
ONE collective on the whole path -- an all-gather of the per-image container byte counts so that
every rank can report the global bpsp (the reference has no distributed code
at all, README.md:86-87).  The conv/entropy path itself needs no exchange: every image is an
independent unit with its own bitstream."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK/WORLD_SIZE/MASTER_* (torchrun).  Returns
    (rank, world, local_rank); a single process without those variables is (0, 1, 0)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29511')
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device('cuda', local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank


def shard_bounds(n_items, rank, world):
    """Contiguous block [lo, hi) of rank: the first (n % world) ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_byte_counts(local_counts, n_items, rank, world, device=None):
    """all-gather of int64 byte counts -> list of n_items counts in global image order."""
    if world == 1:
        return [int(c) for c in local_counts]
    per = max(shard_bounds(n_items, r, world)[1] - shard_bounds(n_items, r, world)[0]
              for r in range(world))
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device()) \
            if dist.get_backend() == 'nccl' else torch.device('cpu')
    mine = torch.full((per,), -1, dtype=torch.int64, device=device)
    if len(local_counts):
        mine[:len(local_counts)] = torch.tensor(list(local_counts), dtype=torch.int64, device=device)
    gathered = torch.empty(world * per, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(gathered, mine)
    g = gathered.cpu().reshape(world, per)
    out = []
    for r in range(world):
        lo, hi = shard_bounds(n_items, r, world)
        out.extend(int(x) for x in g[r, :hi - lo])
    return out


def global_bpsp(byte_counts, subpixels_per_image):
    return sum(byte_counts) * 8.0 / (len(byte_counts) * subpixels_per_image)
