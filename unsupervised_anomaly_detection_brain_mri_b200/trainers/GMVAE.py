"""GMVAE trainer (mirror of reference trainers/GMVAE.py): reconstruction + conditional-prior + w-prior + cluster-prior loss (:58-88),
the AE train / validation loop, and the iterative MAP restoration at test time (:166-197) as a device-resident loop
(``ConvAutoencoderEngine.restore``, shared with VAE_You).  restore_steps == 0 returns the plain reconstruction xz_mu (:170-177).

STATUS: the engine's call sequences (train step, restoration iteration) are verified on CPU against oracle/gmvae_cpu.py through the ABI
emulator, the latent kernel's arithmetic against float64 autograd via a host build of the same header; first hardware run pending
(tests/test_gpu_gmvae.py, opt-in)."""
from .AEMODEL import AEMODEL, Phase, indicate_early_stopping, update_log_dicts  # noqa: F401
from .VAE_You import VAE_You


class GMVAE(VAE_You):
    class Config(AEMODEL.Config):
        def __init__(self):
            super().__init__('GMVAE')
            self.dim_c = 6
            self.dim_z = 1
            self.dim_w = 1
            self.c_lambda = 1
            self.restore_lr = 1e-3
            self.restore_steps = 150
            self.tv_lambda = 1.8

    def __init__(self, sess, config=None, network=None):
        AEMODEL.__init__(self, sess, config, network)
        cfg = self.config
        self.dim_c, self.dim_z, self.dim_w, self.c_lambda = cfg.dim_c, cfg.dim_z, cfg.dim_w, cfg.c_lambda
        self.restore_lr, self.restore_steps, self.tv_lambda_value = cfg.restore_lr, cfg.restore_steps, cfg.tv_lambda
        self.xz_mu = self.outputs['xz_mu']
        self.pc = self.outputs['pc']

    REC_KEY = 'xz_mu'

    def _engine_extra(self):
        cfg = self.config
        return dict(dim_w=int(cfg.dim_w), dim_c=int(cfg.dim_c), c_lambda=float(cfg.c_lambda))

    def reconstruct(self, x, dropout=False):
        if int(self.restore_steps) == 0:
            return AEMODEL.reconstruct(self, x, dropout)
        return VAE_You.reconstruct(self, x, dropout)
