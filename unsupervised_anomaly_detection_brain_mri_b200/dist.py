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


class _DeviceArray:
    """A raw device allocation seen through __cuda_array_interface__ (torch.as_tensor maps it without a copy)."""

    def __init__(self, ptr, numel):
        self.__cuda_array_interface__ = {'shape': (int(numel),), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 3,
                                         'strides': None}


class PeerOptimizer:
    """The optimiser step of a data-parallel run as ONE kernel over NVLink peer memory (csrc/uad_peer.cu: reduce-scatter of the
    gradients, TF-Adam on the own shard, all-gather of the new parameters) instead of `all_reduce` + Adam.

    Moves the flat parameter and gradient buffers of ``fp`` (engine.FlatParams) into a cudaMalloc'ed region whose IPC handle the
    ranks exchange through the process group; ``fp.params`` / ``fp.grads`` become views of that region, so the engine's kernels
    write their gradients where the peers read them and read their parameters where the peers write them.  Must be created before
    the first train step (a captured CUDA graph holds the buffer addresses)."""

    def __init__(self, fp, device):
        import ctypes as C
        from . import abi
        assert dist.is_initialized() and world_size() > 1, 'PeerOptimizer needs an initialised process group'
        self.rank, self.world, self.numel = rank(), world_size(), int(fp.numel)
        L = abi.lib()
        nbytes = L.uad_peer_region_bytes(self.numel)
        region = C.c_void_p()
        abi.call('uad_peer_alloc', nbytes, C.byref(region))
        self.region = region.value
        half = (self.numel * 4 + 255) & ~255
        self._holders = (_DeviceArray(self.region, self.numel), _DeviceArray(self.region + half, self.numel))
        params = torch.as_tensor(self._holders[0], device=device)
        grads = torch.as_tensor(self._holders[1], device=device)
        params.copy_(fp.params)
        grads.zero_()
        fp.params, fp.grads = params, grads
        handle = C.create_string_buffer(64)
        abi.call('uad_peer_ipc_handle', self.region, handle)
        blobs = [None] * self.world
        dist.all_gather_object(blobs, handle.raw)
        self.regions = (C.c_void_p * 16)()
        self._opened = []
        for j, blob in enumerate(blobs):
            if j == self.rank:
                self.regions[j] = self.region
            else:
                out = C.c_void_p()
                abi.call('uad_peer_ipc_open', C.create_string_buffer(blob, 64), C.byref(out))
                self.regions[j] = out.value
                self._opened.append(out.value)
        torch.cuda.synchronize(device)
        dist.barrier()

    def step(self, m, v, lr, beta1, beta2, eps, grad_scale, step_dev, stream, lo=0, hi=None):
        """One optimiser step on the slice [lo, hi) of the flat index space (default: everything); m / v are the FULL moment buffers."""
        from . import abi
        hi = self.numel if hi is None else hi
        abi.call('uad_peer_adam_step', self.regions, self.rank, self.world, self.numel, int(lo), int(hi - lo), m[lo:].data_ptr(),
                 v[lo:].data_ptr(), float(lr), float(beta1), float(beta2), float(eps), float(grad_scale), step_dev.data_ptr(), stream)

    def shard_range(self, lo=0, hi=None):
        hi = self.numel if hi is None else hi
        chunk = (-(-(hi - lo) // self.world) + 3) & ~3
        a = lo + self.rank * chunk
        return a, min(hi, a + chunk)

    def gather_adam_state(self, m, v):
        """Every rank holds the Adam moments of its own shard only; before they are written to a checkpoint, collect them."""
        lo, hi = self.shard_range()
        for buf in (m, v):
            own = torch.zeros_like(buf)
            own[lo:hi] = buf[lo:hi]
            dist.all_reduce(own, op=dist.ReduceOp.SUM)
            buf.copy_(own)
