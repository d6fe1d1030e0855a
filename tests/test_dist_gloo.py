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


def _fanogan_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from oracle import fanogan_cpu as FO
    from oracle import tf_graph_cpu as O
    from unsupervised_anomaly_detection_brain_mri_b200 import dist as udist
    from unsupervised_anomaly_detection_brain_mri_b200 import engine as E
    from unsupervised_anomaly_detection_brain_mri_b200 import fanogan_engine as FE
    udist.init_from_env('gloo')
    S, B = 32, 4
    P = FO.perturb(FO.init_params(S))
    rng = np.random.default_rng(0)
    x = O.synthetic_slices(B, S)
    z = rng.standard_normal((B, 128)).astype(np.float32)
    alpha = rng.random((B, 1), dtype=np.float32)
    fp = E.FlatParams(FE.param_specs(S), 'cpu')
    if rank == 0:
        fp.load(P)
    udist.broadcast_(fp.params)
    errs = []
    for which, scope in (('disc', 'Discriminator'), ('gen', 'Generator'), ('enc', 'Encoder')):
        L = FO.as_leaves(fp.to_numpy(), torch.float64)
        out = FO.wgan_graph(L, udist.shard(x), udist.shard(z), udist.shard(alpha), training=False, dtype=torch.float64, want=(which,))
        G = FO.scope_grads(L, out[{'disc': 'disc_loss', 'gen': 'gen_loss', 'enc': 'enc_loss'}[which]], scope)
        fp.grads.zero_()
        for k, g in G.items():
            fp.g(k).copy_(g.reshape(-1).float())
        lo, hi = fp.subset_ranges(scope + '/')
        udist.allreduce_sum_(fp.grads[lo:hi])                     # the ONLY collective of a train op: its scope's slice
        avg = {k: fp.g(k).numpy() / world for k in G}
        if rank == 0:
            Lf = FO.as_leaves(P, torch.float64)
            of = FO.wgan_graph(Lf, x, z, alpha, training=False, dtype=torch.float64, want=(which,))
            Gf = FO.scope_grads(Lf, of[{'disc': 'disc_loss', 'gen': 'gen_loss', 'enc': 'enc_loss'}[which]], scope)
            gmax = max(float(g.abs().max()) for g in Gf.values())
            errs.append(max(float(np.abs(avg[k] - Gf[k].numpy().reshape(-1)).max()) for k in Gf) / gmax)
    if rank == 0:
        q.put(errs)
    torch.distributed.destroy_process_group()


def test_fanogan_scope_slices_and_two_rank_allreduce():
    """f-AnoGAN: the flat buffer is scope-contiguous (Encoder | Generator | Discriminator), each train op all-reduces only its
    scope's slice, and the rank-averaged local gradients equal the global-batch gradients (losses are batch means; the
    gradient penalty is a mean over (b, w))."""
    from unsupervised_anomaly_detection_brain_mri_b200 import engine as E
    from unsupervised_anomaly_detection_brain_mri_b200 import fanogan_engine as FE
    fp = E.FlatParams(FE.param_specs(64), 'cpu')
    r = [fp.subset_ranges(s + '/') for s in ('Encoder', 'Generator', 'Discriminator')]
    assert r[0][0] == 0 and r[0][1] == r[1][0] and r[1][1] == r[2][0] and r[2][1] == fp.numel
    for (lo, hi), s in zip(r, ('Encoder', 'Generator', 'Discriminator')):
        assert all((lo <= fp.offsets[k] < hi) == k.startswith(s + '/') for k in fp.specs)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fanogan_worker, args=(rk, 2, port, q)) for rk in range(2)]
    for p in procs:
        p.start()
    errs = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert max(errs) < 1e-5, errs
