"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink; gloo for CPU tests).

The hot path shards by slice: each rank runs the same step on its own mini-batch shard and the ONLY data-path
collective is one sum-all-reduce of the flat gradient buffer (8.8 MB fp32) per optimiser step; the 1/world factor
is folded into the fused Adam kernel (grad_scale).  BN is frozen, so there are no cross-sample statistics."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's RANK / WORLD_SIZE / MASTER_* variables (idempotent)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1 and not dist.is_initialized():
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        kw = {}
        if backend == 'nccl':
            kw['device_id'] = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
        dist.init_process_group(backend, **kw)
    return rank(), world_size()


def rank():
    return dist.get_rank() if dist.is_initialized() else 0


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def allreduce_sum_(flat):
    if world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def allreduce_sum_async(flat):
    """Sum-all-reduce that does not block the calling stream: returns the work handle (``.wait()`` makes the CURRENT stream wait
    for the result), or None in a single-process run.  The engine uses it for the decoder's gradient bucket, which is complete
    while the encoder's backward pass is still running."""
    if world_size() > 1:
        return dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
    return None


allreduce_sum_.async_ = allreduce_sum_async


def broadcast_(flat, src=0):
    if world_size() > 1:
        dist.broadcast(flat, src=src)
    return flat


def shard(batch, r=None, w=None):
    """Rank r's contiguous shard of a global batch (global batch must divide evenly, like num_batches drops remainders)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    n = batch.shape[0]
    if n % w:
        raise ValueError(f'global batch {n} does not divide over {w} ranks')
    per = n // w
    return batch[r * per:(r + 1) * per]


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def mean_over_ranks(value, device):
    """Mean of a host scalar over the ranks (every rank gets the same number: e.g. the validation loss that decides early stopping)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t /= world_size()
    return float(t.item())
