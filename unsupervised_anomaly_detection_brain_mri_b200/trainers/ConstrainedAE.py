"""ConstrainedAE trainer (mirror of reference trainers/ConstrainedAE.py): loss = mean_b(L2 + rho * Rec_z) with
L2 = mean_hwc (x - x_hat)^2 and Rec_z = mean_j (z - z_rec)^2 (:37-43); 'reconstructionLoss' (L1) is reported only."""
from .AEMODEL import AEMODEL, Phase, indicate_early_stopping, update_log_dicts  # noqa: F401


class ConstrainedAE(AEMODEL):
    class Config(AEMODEL.Config):
        def __init__(self):
            super().__init__('ConstrainedAE')
            self.rho = 1

    def __init__(self, sess, config=None, network=None):
        super().__init__(sess, config, network)
        self.z = self.outputs['z']
        self.z_rec = self.outputs['z_rec']
        self.rho = self.config.rho
        self.engine.rho = float(self.rho)

    def _eval_engine(self, n):
        eng = super()._eval_engine(n)
        eng.rho = float(self.rho)
        return eng
