"""ceVAE trainer (mirror of reference trainers/ceVAE.py): two-branch L1 + KL (:38-50), input-gradient anomaly (:51),
``reconstruct`` with the gradient-based restoration ``x - lambda * anomaly`` (:119-144)."""
import numpy as np

from .AEMODEL import AEMODEL, Phase, indicate_early_stopping, update_log_dicts  # noqa: F401
from .CE import retrieve_masked_batch


class ceVAE(AEMODEL):
    TWO_INPUTS = True

    class Config(AEMODEL.Config):
        def __init__(self):
            super().__init__('ceVAE')
            self.use_gradient_based_restoration = True

    def __init__(self, sess, config=None, network=None):
        super().__init__(sess, config, network)
        self.reconstruction_ce = self.outputs['x_hat_ce']
        self.z_mu = self.outputs['z_mu']
        self.z_sigma = self.outputs['z_sigma']

    def _make_ce_batch(self, batch, brainmasks, phase):
        # ceVAE.py:105: masked batch only in TRAIN, the plain batch otherwise
        return retrieve_masked_batch(batch, brainmasks) if phase == Phase.TRAIN else batch

    def run_batch(self, batch, phase, batch_ce=None, fetch_maps=False, want_anomaly=True, prefetch=None, prefetch_ce=None):
        if batch_ce is None:
            batch_ce = batch
        if phase != Phase.TRAIN:
            eng = self.engine
            eng.set_inputs(self._feed('x', batch), self._feed('x_ce', batch_ce))
            eng.draw_noise(False, 0.0)
            eng.forward(training=False, dropout_rate=0.0)
            return {k: np.float32(v) for k, v in eng.losses().items()}
        run = super().run_batch(batch, phase, batch_ce=batch_ce, fetch_maps=fetch_maps, want_anomaly=True, prefetch=prefetch,
                                prefetch_ce=prefetch_ce)
        if fetch_maps:
            run['reconstruction_ce'] = self.engine.br[1].xhat.cpu().numpy()
            run['anomaly'] = self.engine.anomaly.cpu().numpy()
        return run

    def reconstruct(self, x, dropout=False):
        """Reference trainers/ceVAE.py:119-144: forward with ``x_ce = x``, every loss tensor fetched; with a truthy
        ``config.use_gradient_based_restoration`` the returned 'reconstruction' is ``x - lambda * anomaly`` (:136-139).

        The reference is called with ONE slice per ``sess.run`` (utils/Evaluation.py:246-253), so the batch mean in
        ``loss_vae`` is a mean over one sample: a stack [N,H,W,C] batched on the device here gets the per-slice gradient
        (``engine.anomaly_per_sample``), i.e. exactly what N single-slice calls return."""
        if x.ndim < 4:
            x = np.expand_dims(x, 0)
        x = np.ascontiguousarray(x, np.float32)
        N = x.shape[0]
        chunk = min(N, int(getattr(self.config, 'evalBatchsize', 128)))
        eng = self._eval_engine(chunk)
        rec = np.empty_like(x)
        anomaly = np.empty_like(x)
        l1_vae = np.empty_like(x)
        rate = self.config.dropout_rate if dropout else 0.0
        for i in range(0, N, chunk):
            xb = x[i:i + chunk]
            n = xb.shape[0]
            if n < chunk:
                xb = np.concatenate([xb, np.zeros((chunk - n,) + xb.shape[1:], np.float32)], 0)
            eng.set_inputs(xb, xb)
            eng._keep = 1.0 / (1.0 - rate) if rate > 0 else 1.0
            eng.draw_noise(dropout, rate)
            eng.forward(training=False, dropout_rate=rate, branches=[0], need_l1=True)
            eng.anomaly_per_sample()
            rec[i:i + n] = eng.br[0].xhat.cpu().numpy()[:n]
            anomaly[i:i + n] = eng.anomaly.cpu().numpy()[:n]
            l1_vae[i:i + n] = eng.br[0].l1.cpu().numpy()[:n]
        results = {'reconstruction': rec, 'anomaly': anomaly, 'L1_vae': l1_vae}
        lam = getattr(self.config, 'use_gradient_based_restoration', False)
        if lam:
            # "not the real 'reconstruction'" (reference comment): the residual x - reconstruction becomes lambda * anomaly
            results['reconstruction'] = x - np.float32(lam) * anomaly
        results['l1err'] = np.sum(np.abs(x - results['reconstruction']))
        results['l2err'] = np.sum(np.sqrt((x - results['reconstruction']) ** 2))
        return results
