"""torchrun worker: N-rank data-parallel step == 1-rank step on the concatenated batch (SURVEY 4 'distributed tests').
Each rank runs the engine on its shard with the NCCL all-reduce hook; rank 0 then runs the whole batch alone."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsupervised_anomaly_detection_brain_mri_b200 import dist as udist  # noqa: E402
from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import make_volume  # noqa: E402
from unsupervised_anomaly_detection_brain_mri_b200.engine import ConvAutoencoderEngine, glorot_init  # noqa: E402


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    rank, world = udist.init_from_env('nccl')
    arch, S, Bg, lr = 'variational_autoencoder', 64, 8, 1e-3
    x = make_volume(S, Bg, seed=5, lesions=False)[0][..., None]
    rng = np.random.default_rng(0)
    eps = rng.standard_normal((Bg, 128)).astype(np.float32)
    masks = {'mu': (rng.uniform(size=(Bg, 128)) >= 0.2).astype(np.float32), 'ls': (rng.uniform(size=(Bg, 128)) >= 0.2).astype(np.float32),
             'dec': (rng.uniform(size=(Bg, 512 if S == 32 else 1024)) >= 0.2).astype(np.float32)}
    eng = ConvAutoencoderEngine(arch, S, batch=Bg // world, device=f'cuda:{local}', seed=3)
    masks['dec'] = masks['dec'][:, :eng.flat]
    udist.broadcast_(eng.fp.params)
    eng.set_inputs(udist.shard(x))
    eng.set_noise(udist.shard(eps), {k: udist.shard(v) for k, v in masks.items()})
    eng.train_step(lr, dropout_rate=0.2, dropout=True, parity_noise=True, allreduce=udist.allreduce_sum_, world=world)
    torch.cuda.synchronize()
    if rank == 0:
        ref = ConvAutoencoderEngine(arch, S, batch=Bg, device=f'cuda:{local}', seed=3)
        ref.set_inputs(x)
        ref.set_noise(eps, masks)
        ref.train_step(lr, dropout_rate=0.2, dropout=True, parity_noise=True)
        torch.cuda.synchronize()
        g_dp = eng.fp.grads.cpu().numpy() / world
        g_1 = ref.fp.grads.cpu().numpy()
        gerr = float(np.abs(g_dp - g_1).max() / np.abs(g_1).max())
        w_dp, w_1 = eng.fp.params.cpu().numpy(), ref.fp.params.cpu().numpy()
        werr = float(np.abs(w_dp - w_1).max())
        frac = float((np.abs(w_dp - w_1) > 1e-3 * lr).mean())
        print(f'DP_EQUIV world={world} grad_rel_err={gerr:.3e} max_weight_diff={werr:.3e} frac_diff={frac:.3e}', flush=True)
        assert gerr < 1e-4 and werr <= 2.001 * lr and frac < 5e-3
    torch.distributed.barrier()
    graph_part(rank, world, local)
    peer_part(rank, world, local)
    fanogan_part(rank, world, local)
    fanogan_peer_part(rank, world, local)
    scoring_part(rank, world, local)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def graph_part(rank, world, local):
    """CUDA-graph replay of the data-parallel step in its three forms - (a) the default: decoder gradient bucket all-reduced on the
    side while the encoder's backward runs, the rest at the end, Adam, all INSIDE the captured step; (b) one all-reduce inside the
    graph (UAD_DP_BUCKETS=0 UAD_GRAPH_ALLREDUCE=1); (c) the collective issued behind the replay (both 0) - gives bit-identical
    weights after five steps (device RNG, same seeds)."""
    arch, S, B, lr = 'variational_autoencoder', 64, 4, 1e-3
    x = udist.shard(make_volume(S, B * world, seed=9, lesions=False)[0][..., None])
    res = []
    for buckets, inside in (('1', '0'), ('0', '1'), ('0', '0')):
        os.environ['UAD_DP_BUCKETS'] = buckets
        os.environ['UAD_GRAPH_ALLREDUCE'] = inside
        eng = ConvAutoencoderEngine(arch, S, batch=B, device=f'cuda:{local}', seed=3)
        udist.broadcast_(eng.fp.params)
        eng.set_inputs(x)
        for _ in range(5):
            eng.train_step(lr, dropout_rate=0.2, dropout=True, allreduce=udist.allreduce_sum_, world=world, use_graph=True)
        torch.cuda.synchronize()
        assert eng.graph is not None and eng._graph_has_update == (inside == '1' or buckets == '1')
        assert (eng._bucket_async is not None) == (buckets == '1')
        res.append(eng.fp.params.clone())
    os.environ.pop('UAD_GRAPH_ALLREDUCE')
    os.environ.pop('UAD_DP_BUCKETS')
    same = bool(torch.equal(res[0], res[1])) and bool(torch.equal(res[1], res[2]))
    t = torch.tensor([int(same)], device=f'cuda:{local}')
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN)
    if rank == 0:
        print(f'DP_EQUIV_GRAPH world={world} bucketed in-graph == single in-graph == outside the graph: {bool(t.item())}', flush=True)
    assert t.item() == 1


def peer_part(rank, world, local):
    """The fused peer-memory optimiser step (csrc/uad_peer.cu: reduce-scatter + Adam + all-gather in one kernel, dist.PeerOptimizer)
    against NCCL all-reduce + uad_adam_tf_step: eager and CUDA-graph replay, six steps.  Two ranks: a + b is commutative, so the
    weights must be BIT-identical; more ranks: the summation order differs from NCCL's, equal within rounding (and identical on
    every rank either way).  The Adam moments of a rank's own shard must equal the NCCL run's."""
    arch, S, B, lr = 'variational_autoencoder', 64, 4, 1e-3
    x = udist.shard(make_volume(S, B * world, seed=13, lesions=False)[0][..., None])
    res = []
    for peer in (False, True):
        os.environ['UAD_DP_BUCKETS'] = '0'
        eng = ConvAutoencoderEngine(arch, S, batch=B, device=f'cuda:{local}', seed=3)
        udist.broadcast_(eng.fp.params)
        if peer:
            eng.enable_peer_optimizer()
            assert eng.peer is not None
        eng.set_inputs(x)
        eng.train_step(lr, dropout_rate=0.2, dropout=True, allreduce=udist.allreduce_sum_, world=world, use_graph=False)
        for _ in range(5):
            eng.train_step(lr, dropout_rate=0.2, dropout=True, allreduce=udist.allreduce_sum_, world=world, use_graph=True)
        torch.cuda.synchronize()
        assert eng.graph is not None
        res.append((eng.fp.params.clone(), eng.fp.m.clone(), eng.fp.v.clone(), eng))
    os.environ.pop('UAD_DP_BUCKETS')
    p_nccl, p_peer = res[0][0], res[1][0]
    diff = float((p_nccl - p_peer).abs().max())
    lo, hi = res[1][3].peer.shard_range()
    m_ok = bool(torch.equal(res[0][1][lo:hi], res[1][1][lo:hi])) if world == 2 else True
    # every rank must hold the same parameters after the all-gather
    ref = p_peer.clone()
    torch.distributed.broadcast(ref, src=0)
    same_everywhere = bool(torch.equal(ref, p_peer))
    ok = (diff == 0.0 if world == 2 else diff <= 2.001 * lr) and m_ok and same_everywhere
    t = torch.tensor([int(ok)], device=f'cuda:{local}')
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN)
    if rank == 0:
        print(f'DP_EQUIV_PEER world={world} max |w_peer - w_nccl| = {diff:.3e}, own-shard moments equal: {m_ok}, '
              f'identical on every rank: {same_everywhere}', flush=True)
    assert t.item() == 1


def fanogan_part(rank, world, local):
    """f-AnoGAN critic step (WGAN-GP): each rank on its shard + all-reduce of the Discriminator slice == whole batch."""
    from unsupervised_anomaly_detection_brain_mri_b200.fanogan_engine import FanoganEngine
    S, Bg, lr = 64, 8, 1e-3
    x = make_volume(S, Bg, seed=7, lesions=False)[0][..., None]
    rng = np.random.default_rng(1)
    z = rng.standard_normal((Bg, 128)).astype(np.float32)
    alpha = rng.random(Bg, dtype=np.float32)

    def run(batch, xs, zs, al, hook, w):
        e = FanoganEngine(S, batch=batch, device=f'cuda:{local}', seed=3)
        e.enable_training()
        if hook is not None:
            udist.broadcast_(e.fp.params)
        e.set_inputs(xs)
        e.set_latent(zs)
        e.alpha.copy_(torch.from_numpy(np.ascontiguousarray(al)))
        res = e.step_disc(lr, dropout_rate=0.0, dropout=False, parity_noise=True, allreduce=hook, world=w)
        torch.cuda.synchronize()
        return e, res

    eng, _ = run(Bg // world, udist.shard(x), udist.shard(z), udist.shard(alpha), udist.allreduce_sum_, world)
    if rank == 0:
        ref, _ = run(Bg, x, z, alpha, None, 1)
        lo, hi = ref.fp.subset_ranges('Discriminator/')
        g_dp = eng.fp.grads[lo:hi].cpu().numpy() / world
        g_1 = ref.fp.grads[lo:hi].cpu().numpy()
        gerr = float(np.abs(g_dp - g_1).max() / np.abs(g_1).max())
        same_rest = bool(torch.equal(eng.fp.params[:lo], ref.fp.params[:lo]))
        print(f'DP_EQUIV_FANOGAN world={world} grad_rel_err={gerr:.3e} other_scopes_untouched={same_rest}', flush=True)
        assert gerr < 1e-4 and same_rest
    torch.distributed.barrier()


def fanogan_peer_part(rank, world, local):
    """f-AnoGAN train ops with the fused peer-memory optimiser on the updated scope's slice (three optimisers, three slices, one
    flag region) == NCCL all-reduce of the slice + Adam: generator, critic and encoder steps, eager then CUDA-graph replay."""
    from unsupervised_anomaly_detection_brain_mri_b200.fanogan_engine import FanoganEngine
    S, B, lr = 64, 4, 1e-3
    x = udist.shard(make_volume(S, B * world, seed=17, lesions=False)[0][..., None])
    z = udist.shard(np.random.default_rng(4).standard_normal((B * world, 128)).astype(np.float32))
    res = []
    for peer in (False, True):
        e = FanoganEngine(S, batch=B, device=f'cuda:{local}', seed=3)
        e.enable_training()
        udist.broadcast_(e.fp.params)
        if peer:
            e.enable_peer_optimizer()
        e.set_inputs(x)
        e.set_latent(z)
        for it in range(3):
            kw = dict(dropout_rate=0.1, dropout=True, allreduce=udist.allreduce_sum_, world=world, use_graph=it > 0)
            e.step_gen(lr, **kw)
            e.step_disc(lr, **kw)
            e.step_disc(lr, **kw)
            e.step_enc(lr, **kw)
        torch.cuda.synchronize()
        res.append(e.fp.params.clone())
    diff = float((res[0] - res[1]).abs().max())
    ref = res[1].clone()
    torch.distributed.broadcast(ref, src=0)
    ok = (diff == 0.0 if world == 2 else diff <= 2.001 * lr * 12) and bool(torch.equal(ref, res[1]))
    t = torch.tensor([int(ok)], device=f'cuda:{local}')
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN)
    if rank == 0:
        print(f'DP_EQUIV_FANOGAN_PEER world={world} max |w_peer - w_nccl| = {diff:.3e}', flush=True)
    assert t.item() == 1


def scoring_part(rank, world, local):
    """Volume-sharded Dice counts: int64 all-reduce of (sum P*G, sum P, sum G) == the counts over the whole stack."""
    from unsupervised_anomaly_detection_brain_mri_b200.trainers import Metrics
    rng = np.random.default_rng(2)
    nvol, Z, S = 4, 6, 32
    diffs = rng.random((nvol, Z, S, S), dtype=np.float32)
    labels = rng.uniform(size=diffs.shape) < 0.05
    mine = slice(rank, None, world)
    sc = Metrics.DeviceScorer(diffs[mine].reshape(-1, S, S), labels[mine].reshape(-1, S, S), device=f'cuda:{local}',
                              allreduce=udist.allreduce_sum_)
    best = Metrics.compute_dice_curve_recursive(None, None, granularity=6, scorer=sc)
    # the threshold-free metrics pool the voxels of all ranks (utils/Evaluation.pool_over_ranks): same AUC as one process
    from unsupervised_anomaly_detection_brain_mri_b200.utils import Evaluation
    order = np.concatenate([np.arange(r, nvol, world) for r in range(world)])
    pd_, pl_ = Evaluation.pool_over_ranks(diffs[mine].reshape(-1), labels[mine].reshape(-1).astype(int))
    assert np.array_equal(pd_, diffs[order].reshape(-1)) and np.array_equal(pl_, labels[order].reshape(-1).astype(int))
    auc_pooled = Metrics.compute_roc(pd_, pl_)[0]
    if rank == 0:
        full = Metrics.DeviceScorer(diffs.reshape(-1, S, S), labels.reshape(-1, S, S), device=f'cuda:{local}')
        ref = Metrics.compute_dice_curve_recursive(None, None, granularity=6, scorer=full)
        auc_ref = Metrics.compute_roc(diffs.reshape(-1), labels.reshape(-1).astype(int))[0]
        print(f'DP_EQUIV_SCORING world={world} best={best} ref={ref} pooled AUC {auc_pooled:.6f} == {auc_ref:.6f}', flush=True)
        assert best == ref and abs(auc_pooled - auc_ref) < 1e-12
    torch.distributed.barrier()


if __name__ == '__main__':
    main()
