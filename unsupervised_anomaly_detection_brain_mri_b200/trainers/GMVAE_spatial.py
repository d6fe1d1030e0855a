"""GMVAE_spatial trainer (mirror of reference trainers/GMVAE_spatial.py): the GMVAE loss with every prior term evaluated per spatial
position of the code and summed over the positions (:66-91), same train / validation / restoration flow as trainers/GMVAE.py.

STATUS: CPU-verified through the ABI emulator (tests/test_engine_emulated.py); first hardware run pending (tests/test_gpu_gmvae.py)."""
from .AEMODEL import AEMODEL
from .GMVAE import GMVAE


class GMVAE_spatial(GMVAE):
    class Config(AEMODEL.Config):
        def __init__(self):
            super().__init__('GMVAE_spatial')
            self.dim_c = 6
            self.dim_z = 1
            self.dim_w = 1
            self.c_lambda = 1
            self.restore_lr = 1e-3
            self.restore_steps = 150
            self.tv_lambda = 1.8
