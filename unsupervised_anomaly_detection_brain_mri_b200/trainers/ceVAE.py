"""ceVAE trainer (mirror of reference trainers/ceVAE.py): two-branch L1 + KL (:38-50), input-gradient anomaly (:51).

The reference class defines ``reconstruct`` twice (:119-144 and :146-...); Python keeps the LAST definition, i.e. the
plain forward pass with x_ce = x, so ``use_gradient_based_restoration`` is never consulted there.  Same here."""
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
            import numpy as np
            return {k: np.float32(v) for k, v in eng.losses().items()}
        run = super().run_batch(batch, phase, batch_ce=batch_ce, fetch_maps=fetch_maps, want_anomaly=True, prefetch=prefetch,
                                prefetch_ce=prefetch_ce)
        if fetch_maps:
            run['reconstruction_ce'] = self.engine.br[1].xhat.cpu().numpy()
            run['anomaly'] = self.engine.anomaly.cpu().numpy()
        return run
