"""Constrained adversarial-autoencoder trainer (mirror of reference trainers/ConstrainedAAE.py): the AAE loop (trainers/AAE.py here)
on the constrained graph - loss = mean_b(L2 + rho * Rec_z) with x_hat re-encoded by the same layers (ConstrainedAAE.py:58-61), a
100-50-1 latent critic, and an optim_gen that also updates the 1x1 bottleneck conv and the latent Dense (they live in the
reference's 'Encoder' scope, models/constrained_adversarial_autoencoder.py:13-29).

STATUS: CPU-verified through the ABI emulator (tests/test_engine_emulated.py); first hardware run pending (tests/test_gpu_aae.py)."""
from .AAE import AAE
from .AEMODEL import AEMODEL


class ConstrainedAAE(AAE):
    class Config(AEMODEL.Config):
        def __init__(self):
            super().__init__('ConstrainedAAE')
            self.rho = 1
            self.scale = 10.0

    def _engine_kwargs(self):
        return dict(scale=float(self.config.scale), constrained=True, rho=float(self.config.rho))

    def train(self, dataset):
        self.engine.rho = float(self.config.rho)
        super().train(dataset)
