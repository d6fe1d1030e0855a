"""CPU, world_size 2 (gloo): the data-parallel contract of the step - shard the global batch by rank, all-reduce-sum
the flat gradient buffer, scale by 1/world in the optimiser == the single-process step on the whole batch."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from oracle import tf_graph_cpu as O
    from unsupervised_anomaly_detection_brain_mri_b200 import dist as udist
    from unsupervised_anomaly_detection_brain_mri_b200 import engine as E
    r, w = udist.init_from_env('gloo')
    assert (r, w) == (rank, world)
    arch, S, B = O.VAE, 32, 4
    P = O.perturb_params(O.init_params(arch, S))
    x = O.synthetic_slices(B, S)
    eps = np.random.default_rng(0).standard_normal((B, 128)).astype(np.float32)
    fp = E.FlatParams(E.param_specs(arch, S), 'cpu')
    if rank == 0:
        fp.load(P)
    udist.broadcast_(fp.params)                                   # rank 0's weights everywhere
    xs, es = udist.shard(x), udist.shard(eps)
    assert xs.shape[0] == B // world
    _, L, G = O.loss_and_grads(arch, fp.to_numpy(), xs, eps=es, training=False, dtype=torch.float64)
    fp.load({k: g.numpy() for k, g in G.items()}, buf=fp.grads)   # local mean-loss gradients
    udist.allreduce_sum_(fp.grads)
    avg = {k: v / world for k, v in fp.to_numpy(fp.grads).items()}
    t = udist.max_over_ranks(float(rank + 1), 'cpu')
    if rank == 0:
        _, _, Gfull = O.loss_and_grads(arch, P, x, eps=eps, training=False, dtype=torch.float64)
        err = max(float(np.abs(avg[k] - Gfull[k].numpy()).max() / max(np.abs(Gfull[k].numpy()).max(), 1e-30)) for k in P)
        q.put((err, t))
    torch.distributed.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_global_batch_gradient():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, t = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-5, err
    assert t == 2.0


def test_shard_rejects_uneven_batches():
    import pytest
    from unsupervised_anomaly_detection_brain_mri_b200 import dist as udist
    with pytest.raises(ValueError):
        udist.shard(np.zeros((5, 2)), 0, 2)
    assert udist.shard(np.arange(8).reshape(8, 1), 1, 4).ravel().tolist() == [2, 3]
