"""VAE trainer (mirror of reference trainers/VAE.py): L1 + KL (:36-42)."""
from .AEMODEL import AEMODEL, Phase, indicate_early_stopping, update_log_dicts  # noqa: F401


class VAE(AEMODEL):
    class Config(AEMODEL.Config):
        def __init__(self):
            super().__init__('VAE')

    def __init__(self, sess, config=None, network=None):
        super().__init__(sess, config, network)
        self.z_mu = self.outputs['z_mu']
        self.z_sigma = self.outputs['z_sigma']
